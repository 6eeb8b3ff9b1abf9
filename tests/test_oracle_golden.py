"""CPU: the oracle restatement (oracle/crn_oracle.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference ships no tests of its own)."""
import numpy as np
import pytest
import torch

from common import CONFIGS, load_golden, make_oracle
from oracle import crn_oracle, synth


def test_param_counts_match_readme():
    # README.md:56,58 of the reference: "6.16MB" / "0.81MB" (= millions of parameters)
    from common import STUDENT, TEACHER
    n_t = sum(int(np.prod(s)) for s in synth.crn_param_shapes(**TEACHER).values())
    n_s = sum(int(np.prod(s)) for s in synth.crn_param_shapes(**STUDENT).values())
    assert n_t == 6160922 and n_s == 812314


def test_chunk_grid_and_segmentation_bit_exact():
    g = load_golden("framing")
    for L, gap, n in zip(g["lengths"], g["gaps"], g["n_chunks"]):
        assert crn_oracle.chunk_grid(int(L), 3200) == (int(gap), int(n))
    for L in (1601, 4000):
        ramp = torch.arange(1, L + 1, dtype=torch.float32).reshape(1, 1, L).repeat(2, 3, 1)
        ramp[1] += 100000
        ramp[:, 1] += 0.25
        ramp[:, 2] += 0.5
        seg, gap = crn_oracle.segmentation(ramp, 3200)
        assert np.array_equal(seg.numpy(), g[f"seg_{L}"])
        ola = crn_oracle.over_add(seg[:, 0].reshape(2, -1, 3200), gap)
        assert np.array_equal(ola.numpy(), g[f"ola_{L}"])


def test_decompress_cirm_and_sisnr():
    g = load_golden("framing")
    out = crn_oracle.decompress_cirm(torch.from_numpy(g["cirm_in"]))
    np.testing.assert_allclose(out.numpy(), g["cirm_out"], rtol=1e-6, atol=1e-6)
    mix, src = synth.make_mixture(3, 5000)
    s = crn_oracle.cal_si_snr(torch.from_numpy(mix[:, 0]), torch.from_numpy(src), torch.tensor([5000, 4000, 3000]))
    np.testing.assert_allclose(float(s), float(g["sisnr"]), rtol=1e-5)


@pytest.mark.parametrize("tag", list(CONFIGS))
def test_oracle_matches_reference_outputs(tag):
    g = load_golden(tag)
    oracle, _ = make_oracle(tag)
    seed, B, L = (int(v) for v in g["meta"])
    mix, _ = synth.make_mixture(B, L)
    with torch.no_grad():
        spec = torch.from_numpy(g["spec_chunk1"])
        # STFT restatement vs the reference's torch.stft path on chunk 1
        x = torch.cat([torch.zeros(B, 3, 1600), torch.from_numpy(mix)], dim=-1)
        seg, gap = crn_oracle.segmentation(x, 3200)
        assert gap == int(g["gap"][0]) and seg.shape[0] // B == int(g["n_chunks"][0])
        N = seg.shape[0] // B
        sp = oracle.stft_trans(seg).reshape(B, N, 3, 201, -1, 2)[:, 1]
        np.testing.assert_allclose(sp.numpy(), g["spec_chunk1"], rtol=0, atol=2e-4)
        oracle.reset()
        fwd = oracle.forward(spec)
        np.testing.assert_allclose(fwd.numpy(), g["fwd_chunk1"], rtol=0, atol=5e-4 * np.abs(g["fwd_chunk1"]).max())
        ist = oracle.istft_trans(torch.from_numpy(g["fwd_chunk1"]))
        np.testing.assert_allclose(ist.numpy(), g["istft_chunk1"], rtol=0, atol=1e-5 * max(1.0, np.abs(g["istft_chunk1"]).max()))
        out = oracle.realtime_process(torch.from_numpy(mix))
        peak = np.abs(g["out"]).max()
        assert np.abs(out.numpy() - g["out"]).max() <= 2e-5 * max(1.0, peak)
        if "out_cont" in g:
            mix2, _ = synth.make_mixture(B, L // 2, first_stream=100)
            out2 = oracle.realtime_process(torch.from_numpy(mix2), True)
            assert np.abs(out2.numpy() - g["out_cont"]).max() <= 2e-5 * max(1.0, np.abs(g["out_cont"]).max())


def test_stream_step_equals_realtime_process():
    """The true-streaming step the CUDA path exposes reproduces realtime_process (SURVEY.md section 3.1 probe)."""
    oracle, _ = make_oracle("crn_small")
    mix, _ = synth.make_mixture(2, 4000)
    x = torch.from_numpy(mix)
    with torch.no_grad():
        ref = oracle.realtime_process(x)
        oracle.reset()
        xp = torch.cat([torch.zeros(2, 3, 1600), x], dim=-1)
        seg, gap = crn_oracle.segmentation(xp, 3200)
        N = seg.shape[0] // 2
        seg = seg.reshape(2, N, 3, 3200)
        carry, outs = None, []
        for n in range(N):
            o, carry = oracle.stream_step(seg[:, n], carry)
            outs.append(o)
        y = torch.cat(outs[1:], dim=-1)[:, 1600:1600 + 4000]
    assert np.abs(y.numpy() - ref.numpy()).max() < 1e-5
