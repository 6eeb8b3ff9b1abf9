"""Training micro-step (train.py:195-198): the oracle restatement's autograd gradients against fixtures produced by the
unmodified reference (oracle/make_golden.py train): pred, d loss/d pred, loss scalars and every parameter gradient."""
import os

import numpy as np
import torch

from common import GOLDEN, SMALL, TEACHER, rel_err
from oracle import crn_oracle, synth


def _oracle(cfg, seed):
    w = synth.make_crn_weights(seed=seed, **cfg)
    return crn_oracle.CRNOracle({k: torch.from_numpy(v) for k, v in w.items()}, segment_length=3200, **cfg)


def test_small_train_step_matches_reference():
    g = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    o = _oracle(SMALL, 7)
    mix, src = synth.make_mixture(2, 8000)
    pred, dpred, losses, grads = crn_oracle.train_step_grads(o, mix, src, [8000, 6500], False)
    assert rel_err(pred.numpy(), g["small_pred"]) < 1e-4
    assert np.allclose(losses, g["small_loss"], atol=2e-3)
    assert rel_err(dpred.numpy(), g["small_dpred"]) < 2e-3
    keys = [k[len("small_grad/"):] for k in g.files if k.startswith("small_grad/")]
    assert len(keys) == 102 and set(keys) == set(grads)  # 102 tensors carry a gradient
    for k in keys:
        assert rel_err(grads[k].numpy(), g["small_grad/" + k]) < 5e-3, k
    # flag=True continuation: state carried from the previous piece, no front pad (CRN_ELU.py:474-481)
    mix2, src2 = synth.make_mixture(2, 4800, first_stream=100)
    pred, dpred, losses, grads = crn_oracle.train_step_grads(o, mix2, src2, [4800, 4800], True)
    assert rel_err(pred.numpy(), g["small_cont_pred"]) < 1e-4
    assert np.allclose(losses, g["small_cont_loss"], atol=2e-3)
    for k in keys:
        assert rel_err(grads[k].numpy(), g["small_cont_grad/" + k]) < 5e-3, k


def test_teacher_train_step_matches_reference():
    g = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    o = _oracle(TEACHER, 0)
    mix, src = synth.make_mixture(1, 6400)
    pred, dpred, losses, grads = crn_oracle.train_step_grads(o, mix, src, [6400], False)
    assert np.allclose(losses, g["teacher_loss"], atol=2e-3)
    assert rel_err(dpred.numpy(), g["teacher_dpred"]) < 2e-3
    n = 0
    for k in grads:
        gn = float(g["teacher_gnorm/" + k])
        assert abs(float(grads[k].norm()) - gn) <= 5e-3 * gn + 1e-9, k
        n += grads[k].numel()
    assert n == 6160906
