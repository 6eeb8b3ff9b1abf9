"""The FullSubNet CPU restatement (oracle/fsn_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py fsn).  Known answer: 6,317,515 parameters."""
import os

import numpy as np
import torch

from common import GOLDEN
from oracle import synth
from oracle.fsn_oracle import FSNOracle, unfold

FSN_SMALL = dict(num_freqs=201, num_mics=3, fb_hidden=64, sb_hidden=32, sb_num_neighbors=15, fb_num_neighbors=0,
                 num_layers=2)
FSN_FULL = dict(num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15, fb_num_neighbors=0,
                num_layers=2)


def load(tag):
    z = np.load(os.path.join(GOLDEN, f"{tag}.npz"))
    return {k: z[k] for k in z.files}


def test_param_count_known_answer():
    assert sum(int(np.prod(s)) for s in synth.fsn_param_shapes(**FSN_FULL).values()) == 6317515


def test_unfold_indices_bit_exact():
    g = load("fsn_small")
    x = torch.from_numpy(g["unfold_in"])
    assert np.array_equal(unfold(x, 2).numpy(), g["unfold_out"])
    assert np.array_equal(unfold(x, 0).numpy(), g["unfold0_out"])
    # f=0 -> 15,14,...,1,0,1,...,15 ; f=200 -> 185,...,200,199,...,185
    ramp = torch.arange(201, dtype=torch.float32).reshape(1, 1, 201, 1)
    u = unfold(ramp, 15)[0, :, 0, :, 0].numpy()
    assert list(u[0]) == list(range(15, 0, -1)) + list(range(0, 16))
    assert list(u[200]) == list(range(185, 201)) + list(range(199, 184, -1))


def test_forward_and_realtime_process_match_reference_small():
    g = load("fsn_small")
    o = FSNOracle(synth.make_fsn_weights(seed=11, **FSN_SMALL), **FSN_SMALL)
    x1 = torch.from_numpy(g["x_chunk1"])
    with torch.no_grad():
        o.reset_state(x1.shape[0])
        f1 = o.forward(x1.clone()).numpy()
        f2 = o.forward(x1.clone()).numpy()
        assert np.abs(f1 - g["fwd_chunk1"]).max() < 2e-5
        assert np.abs(f2 - g["fwd_chunk1_again"]).max() < 2e-5
        B, L = int(g["meta"][1]), int(g["meta"][2])
        mix, _ = synth.make_mixture(B, L)
        y, crm = o.realtime_process(torch.from_numpy(mix))
        assert np.abs(crm.numpy() - g["crm"]).max() < 5e-5
        assert np.abs(y.numpy() - g["out"]).max() < 2e-5
        mix2, _ = synth.make_mixture(B, L // 2, first_stream=100)
        y2, _ = o.realtime_process(torch.from_numpy(mix2), flag=True)
        assert np.abs(y2.numpy() - g["out_cont"]).max() < 2e-5


def test_forward_matches_reference_full_config():
    g = load("fsn_full")
    o = FSNOracle(synth.make_fsn_weights(seed=5, **FSN_FULL), **FSN_FULL)
    x1 = torch.from_numpy(g["x_chunk1"])
    with torch.no_grad():
        o.reset_state(1)
        assert np.abs(o.forward(x1.clone()).numpy() - g["fwd_chunk1"]).max() < 2e-5
