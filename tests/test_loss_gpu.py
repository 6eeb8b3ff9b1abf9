"""GPU parity of the forward loss terms (se_cal_si_snr / se_stoi_loss behind utility.cal_si_snr / stoi_loss and
TemporalCRN.compute_loss) against the reference fixtures.  fp32 arithmetic; stated tolerance 2e-4 on the STOI-like
score (a correlation in [-1, 1]) and 2e-3 dB on SI-SNR."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from common import GOLDEN, make_model
from test_loss_oracle_golden import CASES, pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", list(CASES))
def test_loss_terms_match_reference(tag):
    from speech_enhancement_mi_b200 import utility
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    source, pred = pair(tag)
    lens = torch.from_numpy(g[f"{tag}_lens"])
    sisnr = utility.cal_si_snr(pred.cuda(), source.cuda(), lens)
    stoi = utility.stoi_loss(source.cuda(), pred.cuda(), lens)
    assert abs(float(sisnr) - float(g[f"{tag}_sisnr"])) < 2e-3
    assert abs(float(stoi) - float(g[f"{tag}_stoi"])) < 2e-4
    model = make_model("crn_small")
    with contextlib.redirect_stdout(io.StringIO()) as out:
        loss, mae, s = model.compute_loss(source.cuda(), pred.cuda(), lens)
    assert out.getvalue().strip() != ""  # the reference prints sisnr on every call (CRN_ELU.py:530)
    assert np.allclose([float(loss), float(mae), float(s)], g[f"{tag}_loss"], atol=2e-3)


def test_si_snr_without_length_and_identity():
    from speech_enhancement_mi_b200 import utility
    source, pred = pair("b")
    a = float(utility.cal_si_snr(pred.cuda(), source.cuda()))
    from oracle.crn_oracle import cal_si_snr
    assert abs(a - float(cal_si_snr(pred, source))) < 2e-3
    # a very short item scores the constant 0.99 (utility.py:872-874)
    lens = torch.tensor([500, 600])
    assert abs(float(utility.stoi_loss(source.cuda(), pred.cuda(), lens)) + 0.99) < 1e-6
