"""compute_loss terms: the oracle restatement (oracle/crn_oracle.py) against fixtures from the unmodified reference
(utility.stoi_loss / cal_si_snr through CRN_ELU.TemporalCRN.compute_loss, oracle/make_golden.py loss)."""
import os

import numpy as np
import torch

from common import GOLDEN
from oracle import crn_oracle, synth

CASES = {"a": (3, 24000), "b": (2, 6000), "c": (1, 40000)}


def pair(tag):
    B, L = CASES[tag]
    mix, src = synth.make_mixture(B, L)
    source = torch.from_numpy(src)
    return source, 0.8 * source + 0.2 * torch.from_numpy(mix[:, 0])


def test_loss_terms_match_reference():
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    for tag in CASES:
        source, pred = pair(tag)
        lens = torch.from_numpy(g[f"{tag}_lens"])
        loss, mae, sisnr = crn_oracle.compute_loss(source, pred, lens)
        assert abs(float(mae) - float(g[f"{tag}_stoi"])) < 1e-5
        assert abs(float(-sisnr) - float(g[f"{tag}_sisnr"])) < 1e-3
        assert np.allclose([float(loss), float(mae), float(sisnr)], g[f"{tag}_loss"], atol=1e-3)
