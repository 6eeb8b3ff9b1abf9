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
