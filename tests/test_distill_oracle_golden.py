"""The oracle's restatement of the distillation step (feature taps + distillation loss, distillation_crn.py:343-377,
504-566) pinned to tests/golden/distill.npz, which the UNMODIFIED reference DistillationCRN produced
(oracle/make_golden.py distill).  CPU only."""
import numpy as np
import torch

from common import distill_setup, load_golden, rel_err
from oracle import crn_oracle


def tap_sample(x, stride):
    return x.detach().reshape(-1)[::stride].numpy()


def test_oracle_distillation_step_matches_reference():
    g = load_golden("distill")
    teacher, student, connectors, (mix, src, lens) = distill_setup(g)
    _, ft = teacher.realtime_process(torch.from_numpy(mix), False, return_features=True)
    pred, fs = student.realtime_process(torch.from_numpy(mix), False, return_features=True)
    assert rel_err(pred.numpy(), g["pred"]) < 1e-4
    for i in range(5):
        for nm, x in (("ft", ft[i]), ("fs", fs[i])):
            shape = g[f"{nm}{i}_shape"]
            assert tuple(x.shape) == tuple(shape[:4]), (nm, i, x.shape, shape)
            assert rel_err(tap_sample(x, int(shape[4])), g[f"{nm}{i}_sample"]) < 1e-4, (nm, i)
            assert abs(float(x.norm()) / float(g[f"{nm}{i}_norm"]) - 1) < 1e-4
    loss, stoi, sisnr = crn_oracle.compute_loss(torch.from_numpy(src), pred, torch.tensor(lens))
    dl = crn_oracle.distillation_loss(ft, fs, connectors)
    ref = g["loss"]
    assert abs(float(loss + dl) - ref[0]) < 2e-4 * max(1.0, abs(ref[0]))
    assert abs(float(dl) - ref[3]) < 1e-4 * max(1.0, abs(ref[3]))


def test_distillation_module_has_the_reference_checkpoint_surface():
    """predict_distillation.py:33-34 loads a DistillationCRN checkpoint: same state_dict keys and shapes as the reference
    (teacher.*, student.*, connectors.*), and the same parameter aliasing (distillation_crn.py:527-529)."""
    from common import DISTILL_TEACHER
    from speech_enhancement_mi_b200 import distillation_crn
    g = load_golden("distill")
    model = distillation_crn.DistillationCRN(segment_length=3200, dropout=0.0, **DISTILL_TEACHER)
    mine = [f"{k}:{'x'.join(str(d) for d in v.shape)}" for k, v in model.state_dict().items()]
    assert mine == [str(k) for k in g["state_keys"]]
    shared = sum(1 for pt, ps in zip(model.teacher.parameters(), model.student.parameters())
                 if pt.shape == ps.shape and pt.data_ptr() == ps.data_ptr())
    same_shape = sum(1 for pt, ps in zip(model.teacher.parameters(), model.student.parameters()) if pt.shape == ps.shape)
    assert shared == same_shape > 0
    assert model.student.num_channels == [16, 32, 64, 64] and model.student.hidden == 128
    assert [c[0].weight.shape[:2] for c in model.connectors] == [
        (t, s) for t, s in zip(model.teacher.get_channel_num(), model.student.get_channel_num())]
