"""CPU: the C-ABI library builds, loads and exports every symbol include/se_b200.h declares; host-only entry points
agree with the reference fixtures; without a GPU the product path fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from common import REPO, TEACHER, load_golden
from oracle import synth


@pytest.fixture(scope="module")
def native():
    from speech_enhancement_mi_b200 import _native, build
    build.build()
    return _native


def header_symbols():
    text = open(os.path.join(REPO, "include", "se_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(se_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(native):
    syms = header_symbols()
    assert len(syms) >= 20
    handle = C.CDLL(native.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/se_b200.h but not exported"
    assert sorted(native.SIGNATURES) == syms, "ctypes SIGNATURES must cover exactly the header"
    assert b"sm_100a" in native.lib().se_version()


def test_chunk_grid_matches_reference(native):
    g = load_golden("framing")
    for L, gap, n in zip(g["lengths"], g["gaps"], g["n_chunks"]):
        assert native.chunk_grid(int(L), 3200) == (int(gap), int(n))


def test_state_dict_contract():
    """130 keys incl. the net.0 aliases, reference shapes; synthetic weights load strictly."""
    from speech_enhancement_mi_b200.CRN_ELU import TemporalCRN
    m = TemporalCRN(segment_length=3200, **TEACHER)
    sd = m.state_dict()
    assert len(sd) == 130
    assert sum(p.numel() for p in m.parameters()) == 6160922
    shapes = synth.crn_param_shapes(**TEACHER)
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    aliases = [k for k in sd if ".net.0." in k]
    assert len(aliases) == 22
    for k in aliases:
        assert sd[k].data_ptr() == sd[k.replace(".net.0.", ".conv.")].data_ptr()
    w = synth.with_alias_keys(synth.make_crn_weights(seed=0, **TEACHER))
    res = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(native):
    from speech_enhancement_mi_b200.CRN_ELU import TemporalCRN
    m = TemporalCRN(segment_length=3200, **TEACHER)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|CUDA"):
        m.realtime_process(torch.zeros(1, 3, 4000))
