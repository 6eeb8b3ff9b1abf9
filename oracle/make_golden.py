"""Generate tests/golden/*.npz by executing the UNMODIFIED reference under oracle/shim (build container only).

TEST INFRASTRUCTURE.  Run:  python oracle/make_golden.py
Requires /root/reference (absent on the GPU box, which only consumes the committed fixtures).  The reference has no
tests or golden vectors of its own, so these fixtures are the parity pins: outputs of
reference CRN_ELU.TemporalCRN / distillation_crn.TemporalCRN / utility.{segmentation,over_add,decompress_cIRM,
cal_si_snr} on deterministic synthetic weights and mixtures from oracle/synth.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "shim"), "/root/reference", REPO]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchaudio  # noqa: E402

torchaudio.set_audio_backend = lambda *a, **k: None  # API removed in torchaudio 2.x (reference utility.py:475-476)

import CRN_ELU  # noqa: E402  (reference, unmodified)
import distillation_crn  # noqa: E402  (reference, unmodified)
import utility  # noqa: E402  (reference, unmodified)
import fullsubnet  # noqa: E402  (reference, unmodified)

from oracle import synth  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")

TEACHER = dict(num_channels=[16, 32, 64, 128], num_freqs=201, hidden=512, num_layers=2, num_inputs=3, kernel_size=3)
STUDENT = dict(num_channels=[16, 32, 64, 64], num_freqs=201, hidden=128, num_layers=2, num_inputs=3, kernel_size=3)
SMALL = dict(num_channels=[8, 8, 16, 16], num_freqs=201, hidden=32, num_layers=2, num_inputs=3, kernel_size=3)


def load_into(model, weights):
    sd = {k: torch.from_numpy(v) for k, v in synth.with_alias_keys(weights).items()}
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.eval()


def run_model(cls, cfg, seed, B, L, tag, continuation=False):
    weights = synth.make_crn_weights(seed=seed, **cfg)
    model = load_into(cls(segment_length=3200, dropout=0.0, **cfg), weights)
    mix, _ = synth.make_mixture(B, L)
    x = torch.from_numpy(mix)
    res = {}
    with torch.no_grad():
        y = model.realtime_process(x)
        y = y[0] if isinstance(y, tuple) else y
        res["out"] = y.numpy()
        if continuation:
            mix2, _ = synth.make_mixture(B, L // 2, first_stream=100)
            y2 = model.realtime_process(torch.from_numpy(mix2), True)
            y2 = y2[0] if isinstance(y2, tuple) else y2
            res["out_cont"] = y2.numpy()
        # one isolated forward on the first chunk (fresh state): spectrum in -> enhanced spectrum out
        model.reset()
        seg, gap = model.segmentation(torch.cat([torch.zeros(B, 3, 1600), x], dim=-1))
        N = seg.shape[0] // B
        spec = model.stft_trans(seg).reshape(B, N, 3, 201, -1, 2)
        f0 = model.forward(spec[:, 1].contiguous())
        f0 = f0[0] if isinstance(f0, tuple) else f0
        res["spec_chunk1"] = spec[:, 1].numpy()
        res["fwd_chunk1"] = f0.numpy()
        res["istft_chunk1"] = model.istft_trans(f0).numpy()
        res["gap"] = np.array([gap])
        res["n_chunks"] = np.array([N])
    res["meta"] = np.array([seed, B, L])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **res)
    print(tag, {k: v.shape for k, v in res.items()}, "peak", float(np.abs(res["out"]).max()))


def framing():
    """Integer-exact fixtures: segmentation of an index ramp, over_add of chunk ids, gap / N for several lengths."""
    res = {}
    lengths = [1, 1599, 1600, 1601, 3199, 3200, 3201, 4000, 9600, 48000, 65600]
    gaps, ns = [], []
    for L in lengths:
        ramp = torch.arange(1, L + 1, dtype=torch.float32).reshape(1, 1, L).repeat(2, 3, 1)
        ramp[1] += 100000
        ramp[:, 1] += 0.25
        ramp[:, 2] += 0.5
        seg, gap = utility.segmentation(ramp, 3200)
        gaps.append(gap)
        ns.append(seg.shape[0] // 2)
        if L in (1601, 4000):
            res[f"seg_{L}"] = seg.numpy()
            chunks = seg[:, 0].reshape(2, -1, 3200)
            res[f"ola_{L}"] = utility.over_add(chunks, gap).numpy()
    res["lengths"], res["gaps"], res["n_chunks"] = np.array(lengths), np.array(gaps), np.array(ns)
    m = torch.linspace(-12, 12, 4001)
    res["cirm_in"] = m.numpy()
    res["cirm_out"] = utility.decompress_cIRM(m).numpy()
    mix, src = synth.make_mixture(3, 5000)
    est = torch.from_numpy(mix[:, 0])
    res["sisnr"] = utility.cal_si_snr(est, torch.from_numpy(src), torch.tensor([5000, 4000, 3000])).numpy()
    np.savez_compressed(os.path.join(OUT, "framing.npz"), **res)
    print("framing", gaps, ns)


def losses():
    """compute_loss of the reference (CRN_ELU.py:513-535 -> utility.stoi_loss / cal_si_snr) on synthetic pairs."""
    import contextlib
    import io
    model = CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, **SMALL)
    res = {}
    cases = {"a": (3, 24000, [24000, 20000, 9000]), "b": (2, 6000, [6000, 700]), "c": (1, 40000, [40000])}
    for tag, (B, L, lens) in cases.items():
        mix, src = synth.make_mixture(B, L)
        source = torch.from_numpy(src)
        pred = 0.8 * source + 0.2 * torch.from_numpy(mix[:, 0])  # an "enhanced" signal between clean and noisy
        length = torch.tensor(lens)
        with contextlib.redirect_stdout(io.StringIO()):
            loss, mae, sisnr = model.compute_loss(source, pred, length)
        res[f"{tag}_lens"] = np.array(lens)
        res[f"{tag}_stoi"] = np.array(float(utility.stoi_loss(source, pred, length)))
        res[f"{tag}_sisnr"] = utility.cal_si_snr(pred, source, length).numpy()
        res[f"{tag}_loss"] = np.array([float(loss), float(mae), float(sisnr)])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **res)
    print("losses", {k: v for k, v in res.items() if not k.endswith("lens")})


def train_grads():
    """One training micro-step of the UNMODIFIED reference (train.py:195-198): realtime_process -> compute_loss ->
    backward, on seeded weights / mixtures.  Stores pred, d loss / d pred, the three loss scalars and the gradient of
    every parameter (SMALL: full tensors; TEACHER: per-tensor L2 norm and the first 64 entries)."""
    import contextlib
    import io

    def step(model, mix, src, lens, flag):
        model.zero_grad()
        pred = model.realtime_process(torch.from_numpy(mix), flag)
        pred.retain_grad()
        with contextlib.redirect_stdout(io.StringIO()):
            loss, mae, sisnr = model.compute_loss(torch.from_numpy(src), pred, torch.tensor(lens))
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        return pred.detach().numpy(), pred.grad.numpy(), np.array([float(loss), float(mae), float(sisnr)]), grads

    res = {}
    weights = synth.make_crn_weights(seed=7, **SMALL)
    model = load_into(CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, **SMALL), weights).train()
    mix, src = synth.make_mixture(2, 8000)
    pred, dpred, losses_, grads = step(model, mix, src, [8000, 6500], False)
    res["small_pred"], res["small_dpred"], res["small_loss"] = pred, dpred, losses_
    for k, g in grads.items():
        res["small_grad/" + k] = g.numpy()
    mix2, src2 = synth.make_mixture(2, 4800, first_stream=100)
    pred, dpred, losses_, grads = step(model, mix2, src2, [4800, 4800], True)  # flag=True: state carried (data_c.py:61-63)
    res["small_cont_pred"], res["small_cont_loss"] = pred, losses_
    for k, g in grads.items():
        res["small_cont_grad/" + k] = g.numpy()
    res["small_meta"] = np.array([7, 2, 8000, 4800])

    weights = synth.make_crn_weights(seed=0, **TEACHER)
    model = load_into(CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, **TEACHER), weights).train()
    mix, src = synth.make_mixture(1, 6400)
    pred, dpred, losses_, grads = step(model, mix, src, [6400], False)
    res["teacher_pred"], res["teacher_dpred"], res["teacher_loss"] = pred, dpred, losses_
    for k, g in grads.items():
        res["teacher_gnorm/" + k] = np.array(float(g.norm()))
        res["teacher_ghead/" + k] = g.flatten()[:64].numpy()
    res["teacher_meta"] = np.array([0, 1, 6400])
    np.savez_compressed(os.path.join(OUT, "train_grads.npz"), **res)
    print("train_grads", res["small_loss"], res["small_cont_loss"], res["teacher_loss"],
          "n grads", sum(1 for k in res if k.startswith("small_grad/")))


DISTILL_TEACHER = dict(num_channels=[16, 32, 32, 48], num_freqs=201, hidden=64, num_layers=2, num_inputs=3, kernel_size=3)


def tap_sample(x, n=4096):
    """Every stride-th element of the flattened tensor (at most ~n of them): pins values AND layout of a large tap."""
    flat = x.detach().reshape(-1)
    stride = max(1, flat.numel() // n)
    return flat[::stride].numpy().copy(), stride


def distill():
    """One training step of the UNMODIFIED reference DistillationCRN (distillation_crn.py:504-566): teacher and student
    realtime_process with their feature taps, compute_loss + distillation_loss, backward.  The teacher is built without
    `path`, so it is trainable and receives gradients through the margin / target of the distillation loss; the
    student's same-shaped parameters alias the teacher's (distillation_crn.py:527-529).  Weights are loaded teacher
    first, student second (the aliased tensors therefore hold the student's values).  Stores the loss scalars, pred,
    norm + strided samples of every tap and of d loss / d student tap, norm + first 64 entries of every parameter
    gradient, and the connector parameters (torch.manual_seed(0) initialisation) with their full gradients."""
    import contextlib
    import io
    torch.manual_seed(0)
    model = distillation_crn.DistillationCRN(segment_length=3200, dropout=0.0, **DISTILL_TEACHER)
    wt = synth.make_crn_weights(seed=21, **DISTILL_TEACHER)
    ws = synth.make_crn_weights(seed=22, **STUDENT)
    model.teacher.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(wt).items()}, strict=True)
    model.student.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(ws).items()}, strict=True)
    model.train()
    res = {}
    for k, v in model.connectors.state_dict().items():
        res["connector/" + k] = v.detach().numpy().copy()
    # the checkpoint surface predict_distillation.py:33-34 loads: every state_dict key with its shape
    res["state_keys"] = np.array([f"{k}:{'x'.join(str(d) for d in v.shape)}" for k, v in model.state_dict().items()])

    # the same computation as DistillationCRN.forward (distillation_crn.py:560-565), unrolled only to keep the taps
    def step(mix, src, lens, flag, tag):
        model.zero_grad()
        noisy, clean, length = torch.from_numpy(mix), torch.from_numpy(src), torch.tensor(lens)
        _, ft = model.teacher.realtime_process(noisy, flag)
        pred, fs = model.student.realtime_process(noisy, flag)
        for f in fs:
            f.retain_grad()
        with contextlib.redirect_stdout(io.StringIO()):
            loss, stoi, sisnr = model.student.compute_loss(clean, pred, length)
        dl = model.distillation_loss(ft, fs)
        loss = loss + dl
        loss.backward()
        res[tag + "loss"] = np.array([float(loss), float(stoi), float(sisnr), float(dl)])
        res[tag + "pred"] = pred.detach().numpy()
        for i in range(len(ft)):
            for nm, x in (("ft", ft[i]), ("fs", fs[i]), ("dfs", fs[i].grad)):
                smp, stride = tap_sample(x)
                res[f"{tag}{nm}{i}_sample"] = smp
                res[f"{tag}{nm}{i}_norm"] = np.array(float(x.detach().norm()))
                res[f"{tag}{nm}{i}_shape"] = np.array(list(x.shape) + [stride])
        for who in ("teacher", "student"):
            for k, p in getattr(model, who).named_parameters():
                if p.grad is not None:
                    res[f"{tag}{who}_gnorm/{k}"] = np.array(float(p.grad.norm()))
                    res[f"{tag}{who}_ghead/{k}"] = p.grad.flatten()[:64].numpy().copy()
        for k, p in model.connectors.named_parameters():
            res[f"{tag}connector_grad/{k}"] = p.grad.numpy().copy()

    mix, src = synth.make_mixture(2, 4000)
    step(mix, src, [4000, 3300], False, "")
    mix2, src2 = synth.make_mixture(2, 3200, first_stream=100)
    step(mix2, src2, [3200, 3200], True, "cont_")
    # the forward() wrapper itself agrees with the unrolled step (no state is touched: flag=False resets)
    with contextlib.redirect_stdout(io.StringIO()):
        l2, s2, n2 = model(torch.from_numpy(mix), torch.from_numpy(src), torch.tensor([4000, 3300]), False)
    assert abs(float(l2) - float(res["loss"][0])) < 1e-5 * max(1.0, abs(float(l2))), (float(l2), res["loss"])
    res["meta"] = np.array([21, 22, 2, 4000, 3200])
    np.savez_compressed(os.path.join(OUT, "distill.npz"), **res)
    print("distill", res["loss"], res["cont_loss"], "taps", [tuple(res[f"fs{i}_shape"]) for i in range(5)])


FSN_SMALL = dict(num_freqs=201, num_mics=3, fb_hidden=64, sb_hidden=32, sb_num_neighbors=15, fb_num_neighbors=0,
                 num_layers=2)
FSN_FULL = dict(num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15, fb_num_neighbors=0,
                num_layers=2)


def run_fsn(cfg, seed, B, L, tag, continuation=False):
    """Reference FullSubNet.realtime_process(train=False) (fullsubnet.py:903-961) + one isolated forward + unfold."""
    weights = synth.make_fsn_weights(seed=seed, **cfg)
    model = fullsubnet.FullSubNet(
        num_freqs=cfg["num_freqs"], look_ahead=0, sequence_model="LSTM", fb_num_neighbors=cfg["fb_num_neighbors"],
        sb_num_neighbors=cfg["sb_num_neighbors"], fb_output_activate_function="ReLU", sb_output_activate_function=False,
        fb_model_hidden_size=cfg["fb_hidden"], sb_model_hidden_size=cfg["sb_hidden"], num_mics=cfg["num_mics"],
        norm_type="offline_laplace_norm", num_groups_in_drop_band=2, num_layers=cfg["num_layers"], weight_init=False,
        sample_rate=16000, segment_length=3200, win_length=25, hop_length=10, n_fft=400)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    assert not missing and not unexpected
    model.eval()
    mix, src = synth.make_mixture(B, L)
    x = torch.from_numpy(mix)
    s3 = torch.from_numpy(np.repeat(src[:, None, :], 3, axis=1).copy())  # source [B, M, L] (only mic 0 is used)
    res = {}
    with torch.no_grad():
        pred, crm, sf, xf = model.realtime_process(x, s3, flag=False, train=False)
        res["out"], res["crm"], res["sf"], res["xf"] = pred.numpy(), crm.numpy(), sf.numpy(), xf.numpy()
        if continuation:
            mix2, src2 = synth.make_mixture(B, L // 2, first_stream=100)
            s32 = torch.from_numpy(np.repeat(src2[:, None, :], 3, axis=1).copy())
            pred2 = model.realtime_process(torch.from_numpy(mix2), s32, flag=True, train=False)[0]
            res["out_cont"] = pred2.numpy()
        # isolated forward on chunk 1 with fresh state
        seg, gap = model.segmentation(torch.cat([torch.zeros(B, 3, 1600), x], dim=-1))
        spec = model.stft_trans(seg)  # [B*N, 2M, F, T]
        N = spec.shape[0] // B
        x1 = spec.reshape(B, N, 6, 201, -1)[:, 1].contiguous()
        model.reset_state(B, x1.dtype, x1.device)
        res["x_chunk1"] = x1.numpy().copy()
        res["fwd_chunk1"] = model.forward(x1.clone()).numpy()
        res["fwd_chunk1_again"] = model.forward(x1.clone()).numpy()  # second call: running CumLayerNorm mean + LSTM state
        small = torch.arange(2 * 2 * 7 * 3, dtype=torch.float32).reshape(2, 2, 7, 3)
        res["unfold_in"], res["unfold_out"] = small.numpy(), fullsubnet.BaseModel.unfold(small, 2).numpy()
        res["unfold0_out"] = fullsubnet.BaseModel.unfold(small, 0).numpy()
    res["meta"] = np.array([seed, B, L])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **res)
    n = sum(int(np.prod(s)) for s in synth.fsn_param_shapes(**cfg).values())
    print(tag, {k: v.shape for k, v in res.items()}, "peak", float(np.abs(res["out"]).max()), "params", n)


def run_fsn_whole(cfg, seed, B, L, tag):
    """Reference FullSubNet.realtime_process(train=True) (fullsubnet.py:921-927: all chunks concatenated into ONE forward,
    both CumLayerNorms see the whole utterance) plus a flag=True continuation piece, with the 4-tuple outputs."""
    weights = synth.make_fsn_weights(seed=seed, **cfg)
    model = fullsubnet.FullSubNet(
        num_freqs=cfg["num_freqs"], look_ahead=0, sequence_model="LSTM", fb_num_neighbors=cfg["fb_num_neighbors"],
        sb_num_neighbors=cfg["sb_num_neighbors"], fb_output_activate_function="ReLU", sb_output_activate_function=False,
        fb_model_hidden_size=cfg["fb_hidden"], sb_model_hidden_size=cfg["sb_hidden"], num_mics=cfg["num_mics"],
        norm_type="offline_laplace_norm", num_groups_in_drop_band=2, num_layers=cfg["num_layers"], weight_init=False,
        sample_rate=16000, segment_length=3200, win_length=25, hop_length=10, n_fft=400)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    assert not missing and not unexpected
    model.eval()
    mix, src = synth.make_mixture(B, L)
    s3 = torch.from_numpy(np.repeat(src[:, None, :], 3, axis=1).copy())
    res = {}
    with torch.no_grad():
        pred, crm, sf, xf = model.realtime_process(torch.from_numpy(mix), s3, flag=False, train=True)
        res["out"], res["crm"], res["sf"], res["xf"] = pred.numpy(), crm.numpy(), sf.numpy(), xf.numpy()
        mix2, src2 = synth.make_mixture(B, L // 2, first_stream=100)
        s32 = torch.from_numpy(np.repeat(src2[:, None, :], 3, axis=1).copy())
        res["out_cont"] = model.realtime_process(torch.from_numpy(mix2), s32, flag=True, train=True)[0].numpy()
    res["meta"] = np.array([seed, B, L])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **res)
    print(tag, {k: v.shape for k, v in res.items()}, "peak", float(np.abs(res["out"]).max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "fsn_whole" in sys.argv[1:]:
        run_fsn_whole(FSN_SMALL, 11, 2, 4000, "fsn_small_whole")
        sys.exit(0)
    if "loss" in sys.argv[1:]:
        losses()
        sys.exit(0)
    if "train" in sys.argv[1:]:
        train_grads()
        sys.exit(0)
    if "distill" in sys.argv[1:]:
        distill()
        sys.exit(0)
    if "fsn" not in sys.argv[1:]:
        framing()
        losses()
        train_grads()
        distill()
        run_model(CRN_ELU.TemporalCRN, SMALL, 7, 2, 4000, "crn_small", continuation=True)
        run_model(CRN_ELU.TemporalCRN, TEACHER, 0, 2, 8000, "crn_teacher", continuation=True)
        run_model(distillation_crn.TemporalCRN, STUDENT, 3, 2, 8000, "crn_student")
    if "fsn" in sys.argv[1:] or len(sys.argv) == 1:
        run_fsn(FSN_SMALL, 11, 2, 4000, "fsn_small", continuation=True)
        run_fsn(FSN_FULL, 5, 1, 2400, "fsn_full")
    n_t = sum(int(np.prod(s)) for s in synth.crn_param_shapes(**TEACHER).values())
    n_s = sum(int(np.prod(s)) for s in synth.crn_param_shapes(**STUDENT).values())
    print("param counts (README.md:56,58 say 6.16 / 0.81 M):", n_t, n_s)
