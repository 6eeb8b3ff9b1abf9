"""Per-tensor gradient parity table of the training micro-step against tests/golden/train_grads.npz (debug aid).

    python tools/diag_train.py [small|teacher] [fp32|tf32]
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
from common import GOLDEN, make_model, rel_err  # noqa: E402
from oracle import synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "small"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
g = np.load(os.path.join(GOLDEN, "train_grads.npz"))
if which == "small":
    model, (mix, src), lens = make_model("crn_small", precision=prec), synth.make_mixture(2, 8000), [8000, 6500]
else:
    model, (mix, src), lens = make_model("crn_teacher", precision=prec), synth.make_mixture(1, 6400), [6400]
model = model.cuda().train()
pred = model.realtime_process(torch.from_numpy(mix).cuda(), False)
pred.retain_grad()
with contextlib.redirect_stdout(io.StringIO()):
    loss, mae, sisnr = model.compute_loss(torch.from_numpy(src).cuda(), pred, torch.tensor(lens))
loss.backward()
print("pred rel err", rel_err(pred.detach().cpu().numpy(), g[f"{which}_pred"]))
print("loss", float(loss), float(mae), float(sisnr), "ref", g[f"{which}_loss"])
print("dpred rel err", rel_err(pred.grad.cpu().numpy(), g[f"{which}_dpred"]))
for k, p in model.named_parameters():
    if which == "small":
        key = "small_grad/" + k
        if key not in g.files:
            print(f"{k:50s} (no reference gradient) ours max {float(p.grad.abs().max()):.3g}")
            continue
        print(f"{k:50s} rel {rel_err(p.grad.cpu().numpy(), g[key]):.3e}  ref peak {np.abs(g[key]).max():.3e}")
    else:
        key = "teacher_gnorm/" + k
        if key not in g.files:
            continue
        gn = float(g[key])
        print(f"{k:50s} norm ours {float(p.grad.norm()):.4e} ref {gn:.4e}  head err "
              f"{np.abs(p.grad.cpu().numpy().reshape(-1)[:64] - g['teacher_ghead/' + k]).max():.3e}")
