"""Drop-in for the hot-path functions of the reference's ``utility`` module (utility.py:312-403), running in the
sm_100a kernels behind include/se_b200.h.  No PyTorch / CPU compute fallback."""
from __future__ import annotations

import ctypes as C
import os

import torch

from ._native import check, chunk_grid, lib


def _device_of(x):
    if x.is_cuda:
        return x.device.index if x.device.index is not None else torch.cuda.current_device()
    return int(os.environ.get("LOCAL_RANK", "0"))


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def padding(x, K):
    """utility.py:312-336 -- returns (padded signal, gap).  Only the index arithmetic is needed by callers of this
    package (the native segmentation pads implicitly); kept for API parity."""
    B, Cn, L = x.shape
    gap, _ = chunk_grid(L, K)
    P = K // 2
    out = torch.zeros((B, Cn, P + L + gap + P), dtype=x.dtype, device=x.device)
    out[..., P:P + L] = x
    return out, gap


def segmentation(x, K):
    """[B, C, L] -> ([B*N, C, K], gap): 50 %-overlap chunks of the zero-padded signal (utility.py:339-370)."""
    B, Cn, L = x.shape
    dev = _device_of(x)
    gap, N = chunk_grid(L, K)
    with torch.cuda.device(dev):
        xd = x.detach().to(device=f"cuda:{dev}", dtype=torch.float32).contiguous()
        out = torch.empty((B * N, Cn, K), dtype=torch.float32, device=xd.device)
        g, n = C.c_int(0), C.c_int(0)
        check(lib().se_segmentation(xd.data_ptr(), B, Cn, L, K, out.data_ptr(), C.byref(g), C.byref(n), _stream(dev)),
              "se_segmentation")
    return out.to(x.device), g.value


def over_add(x, gap):
    """[C, N, K] -> [C, L]: average of the two chunk tilings, `gap` tail samples dropped (utility.py:373-403)."""
    Cn, N, K = x.shape
    dev = _device_of(x)
    Lout = N * (K // 2) - K // 2 - gap
    with torch.cuda.device(dev):
        xd = x.detach().to(device=f"cuda:{dev}", dtype=torch.float32).contiguous()
        out = torch.empty((Cn, Lout), dtype=torch.float32, device=xd.device)
        check(lib().se_over_add(xd.data_ptr(), Cn, N, K, int(gap), out.data_ptr(), _stream(dev)), "se_over_add")
    return out.to(x.device)


def cal_si_snr(separated, source, length=None, eps=1e-8):
    raise NotImplementedError("cal_si_snr (utility.py:207-223): training-loss kernels are not built yet (DESIGN.md)")


def stoi_loss(source, pred, length):
    raise NotImplementedError("stoi_loss (utility.py:821-916): training-loss kernels are not built yet (DESIGN.md)")
