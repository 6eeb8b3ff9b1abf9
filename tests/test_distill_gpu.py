"""Distillation training step on the CUDA path against tests/golden/distill.npz, produced
by the UNMODIFIED reference DistillationCRN (distillation_crn.py:504-566): feature taps of teacher and student, the
loss, the gradient arriving at every student tap, and the gradients of student, teacher and connector parameters,
for a fresh piece and a flag=True continuation piece.  fp32 mode; tolerances stated per check."""
import contextlib
import io

import numpy as np
import pytest
import torch

from common import DISTILL_TEACHER, STUDENT, load_golden, rel_err
from oracle import synth

pytestmark = pytest.mark.gpu


def build(g, precision="fp32"):
    from speech_enhancement_mi_b200 import distillation_crn
    model = distillation_crn.DistillationCRN(segment_length=3200, dropout=0.0, precision=precision, **DISTILL_TEACHER)
    wt = synth.make_crn_weights(seed=21, **DISTILL_TEACHER)
    ws = synth.make_crn_weights(seed=22, **STUDENT)
    # same order as the fixture: teacher first, student second -- the aliased tensors end up with the student's values
    model.teacher.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(wt).items()}, strict=True)
    model.student.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(ws).items()}, strict=True)
    model.connectors.load_state_dict({k[len("connector/"):]: torch.from_numpy(v) for k, v in g.items()
                                      if k.startswith("connector/")})
    return model.cuda().train()


def step(model, mix, src, lens, flag):
    model.zero_grad()
    noisy, clean = torch.from_numpy(mix).cuda(), torch.from_numpy(src).cuda()
    _, ft = model.teacher.realtime_process(noisy, flag)
    pred, fs = model.student.realtime_process(noisy, flag)
    for f in fs:
        f.retain_grad()
    with contextlib.redirect_stdout(io.StringIO()):
        loss, stoi, sisnr = model.student.compute_loss(clean, pred, torch.tensor(lens))
    dl = model.distillation_loss(ft, fs)
    (loss + dl).backward()
    return pred, ft, fs, np.array([float(loss + dl), float(stoi), float(sisnr), float(dl)])


def check(model, g, tag, pred, ft, fs, losses):
    ref = g[tag + "loss"]
    assert np.abs(losses - ref).max() < 5e-4 * max(1.0, np.abs(ref).max()), (losses, ref)
    assert rel_err(pred.detach().cpu().numpy(), g[tag + "pred"]) < 5e-5
    for i in range(5):
        for nm, x in (("ft", ft[i]), ("fs", fs[i]), ("dfs", fs[i].grad)):
            shape = g[f"{tag}{nm}{i}_shape"]
            assert tuple(x.shape) == tuple(shape[:4]), (nm, i, tuple(x.shape), shape)
            smp = x.detach().reshape(-1)[::int(shape[4])].cpu().numpy()
            # taps: fp32 re-association only.  d loss / d student tap passes through the connector's BatchNorm (batch
            # statistics) and the discontinuous margin mask of distillation_crn.py:561: entries next to the threshold flip
            tol = 2e-3 if nm == "dfs" else 2e-4
            err = rel_err(smp, g[f"{tag}{nm}{i}_sample"])
            assert err < tol, (tag, nm, i, err)
            assert abs(float(x.detach().norm()) / float(g[f"{tag}{nm}{i}_norm"]) - 1) < tol, (tag, nm, i)
    bad = []
    for who in ("teacher", "student"):
        n_checked = 0
        for k, p in getattr(model, who).named_parameters():
            key = f"{tag}{who}_gnorm/{k}"
            if key not in g:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (who, k)
                continue
            n_checked += 1
            gn, head = float(g[key]), g[f"{tag}{who}_ghead/{k}"]
            mine = p.grad.detach().flatten()
            scale = max(gn / max(1.0, mine.numel() ** 0.5), 1e-12)  # rms of the reference gradient
            if abs(float(mine.norm()) - gn) > 1e-3 * gn + 1e-9 or \
                    np.abs(mine[:64].cpu().numpy() - head).max() > 1e-3 * max(np.abs(head).max(), scale):
                bad.append((who, k, float(mine.norm()), gn))
        assert n_checked > 50, (who, n_checked)
    assert not bad, bad
    for k, p in model.connectors.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), g[f"{tag}connector_grad/{k}"]) < 1e-3, k


def test_distillation_step_matches_reference():
    g = load_golden("distill")
    model = build(g)
    mix, src = synth.make_mixture(2, 4000)
    check(model, g, "", *step(model, mix, src, [4000, 3300], False))
    mix2, src2 = synth.make_mixture(2, 3200, first_stream=100)
    check(model, g, "cont_", *step(model, mix2, src2, [3200, 3200], True))
    # the module's own forward (distillation_crn.py:560-565)
    with contextlib.redirect_stdout(io.StringIO()):
        loss, stoi, sisnr = model(torch.from_numpy(mix).cuda(), torch.from_numpy(src).cuda(), torch.tensor([4000, 3300]),
                                  False)
    ref = g["loss"]
    assert abs(float(loss) - ref[0]) < 5e-4 * max(1.0, abs(ref[0]))
    assert abs(float(stoi) - ref[1]) < 5e-4 and abs(float(sisnr) - ref[2]) < 5e-4 * max(1.0, abs(ref[2]))
    loss.backward()


def test_student_eval_serving_path_returns_empty_features():
    """predict_distillation.py:84 discards the features: in eval mode the streaming kernels run and the list is empty;
    return_features=True switches the taps on without autograd (the frozen teacher of DistillationCRN(path=...))."""
    from common import make_model
    model = make_model("crn_student", precision="fp32").cuda()
    mix, _ = synth.make_mixture(1, 4800)
    x = torch.from_numpy(mix).cuda()
    pred, feats = model.realtime_process(x)
    assert feats == []
    model.return_features = True
    with torch.no_grad():
        pred2, feats2 = model.realtime_process(x)
    assert len(feats2) == 5 and [f.shape[1] for f in feats2] == model.get_channel_num()
    assert not feats2[0].requires_grad
    assert rel_err(pred2.cpu().numpy(), pred.cpu().numpy()) < 2e-5
