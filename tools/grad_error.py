#!/usr/bin/env python
"""GPU: worst parameter-gradient error of the fp32 training step against the unmodified reference's gradients
(tests/golden/train_grads.npz, small configuration), for the current environment's kernel switches.

    python tools/grad_error.py            # default kernels
    SE_B200_BWD_MMA=0 SE_B200_GRU_CLUSTER=0 python tools/grad_error.py   # round-1 backward
"""
import contextlib
import io
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from common import GOLDEN, make_model, rel_err  # noqa: E402
from oracle import synth  # noqa: E402

g = np.load(os.path.join(GOLDEN, "train_grads.npz"))
model = make_model("crn_small", precision="fp32").cuda().train()
mix, src = synth.make_mixture(2, 8000)
model.zero_grad()
pred = model.realtime_process(torch.from_numpy(mix).cuda(), False)
pred.retain_grad()
with contextlib.redirect_stdout(io.StringIO()):
    loss, mae, sisnr = model.compute_loss(torch.from_numpy(src).cuda(), pred, torch.tensor([8000, 6500]))
loss.backward()
errs = {}
for k, p in model.named_parameters():
    key = "small_grad/" + k
    if p.grad is not None and key in g.files:
        errs[k] = rel_err(p.grad.detach().cpu().numpy(), g[key])
worst = max(errs, key=errs.get)
print("switches:", {k: v for k, v in os.environ.items() if k.startswith("SE_B200_")})
print(f"pred error {rel_err(pred.detach().cpu().numpy(), g['small_pred']):.3g}; d loss / d pred error "
      f"{rel_err(pred.grad.cpu().numpy(), g['small_dpred']):.3g}")
print(f"worst parameter gradient: {worst} {errs[worst]:.3g} of the tensor's peak; median {np.median(list(errs.values())):.3g}")
