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


def _loss_args(a, b, length):
    dev = _device_of(a if a.is_cuda else b)
    ad = a.detach().to(device=f"cuda:{dev}", dtype=torch.float32).contiguous()
    bd = b.detach().to(device=f"cuda:{dev}", dtype=torch.float32).contiguous()
    if ad.dim() == 3:  # stoi_loss squeezes a trailing singleton (utility.py:845-846)
        ad, bd = ad.squeeze(-1), bd.squeeze(-1)
    B, L = ad.shape
    if length is None:
        ld = None
    else:
        ld = torch.as_tensor(length).to(device=ad.device, dtype=torch.int32).contiguous()
    return dev, ad, bd, ld, B, L


def cal_si_snr(separated, source, length=None, eps=1e-8):
    """Mean SI-SNR in dB over the batch (utility.py:207-223).  Forward only (no autograd graph)."""
    dev, sd, td, ld, B, L = _loss_args(separated, source, length)
    with torch.cuda.device(dev):
        out = torch.empty((), dtype=torch.float32, device=sd.device)
        check(lib().se_cal_si_snr(sd.data_ptr(), td.data_ptr(), None if ld is None else ld.data_ptr(), B, L,
                                  out.data_ptr(), _stream(dev)), "se_cal_si_snr")
    return out.reshape(1)  # the reference accumulates [1]-shaped terms (keepdim sums)


def stoi_loss(y_true_batch, y_pred_batch, lens, reduction="mean"):
    """-mean(STOI-like score) (utility.py:821-916).  Forward only.  The reference moves everything to the CPU and returns
    a CPU scalar; this returns a CUDA scalar."""
    if reduction != "mean":
        raise NotImplementedError("only reduction='mean' (the only use: CRN_ELU.py:526) is built")
    dev, td, pd, ld, B, L = _loss_args(y_true_batch, y_pred_batch, lens)
    with torch.cuda.device(dev):
        out = torch.empty((), dtype=torch.float32, device=td.device)
        check(lib().se_stoi_loss(td.data_ptr(), pd.data_ptr(), ld.data_ptr(), B, L, out.data_ptr(), _stream(dev)),
              "se_stoi_loss")
    return out
