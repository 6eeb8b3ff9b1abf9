"""GPU diagnostic (test infrastructure): per-layer error of the CUDA path vs the oracle on one chunk, then end to end.
Usage: python tools/diag_layers.py [tag] [precision]   -> prints a table; exits non-zero on gross mismatch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from common import load_golden, make_model, make_oracle, rel_err, si_sdr_db  # noqa: E402
from oracle import synth  # noqa: E402
from speech_enhancement_mi_b200._native import check, lib  # noqa: E402


def read(model, name, b):
    buf = np.zeros(4 * 1024 * 1024, dtype=np.float32)
    dims = (C.c_int * 3)()
    check(lib().se_debug_read(model._ctx, name.encode(), b, buf.ctypes.data, buf.size, dims), "se_debug_read")
    t, f, c = dims[0], dims[1], dims[2]
    return buf[: t * f * c].reshape(t, f, c)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "crn_small"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    g = load_golden(tag)
    oracle, _ = make_oracle(tag)
    model = make_model(tag, prec)
    B = g["spec_chunk1"].shape[0]
    spec = torch.from_numpy(g["spec_chunk1"])
    worst = 0.0
    with torch.no_grad():
        # --- pieces ---
        mix, _ = synth.make_mixture(B, int(g["meta"][2]))
        x = torch.cat([torch.zeros(B, 3, 1600), torch.from_numpy(mix)], dim=-1)
        from oracle import crn_oracle
        seg, gap = crn_oracle.segmentation(x, 3200)
        N = seg.shape[0] // B
        seg_d, gap_d = model.segmentation(x.cuda())
        print("segmentation exact:", bool(np.array_equal(seg_d.cpu().numpy(), seg.numpy())), gap == gap_d)
        st = model.stft_trans(seg.cuda()).cpu()
        st_o = oracle.stft_trans(seg)
        print("stft_trans rel err:", rel_err(st.numpy(), st_o.numpy()))
        ist = model.istft_trans(torch.from_numpy(g["fwd_chunk1"]).cuda()).cpu()
        print("istft_trans rel err vs golden:", rel_err(ist.numpy(), g["istft_chunk1"]))
        # --- layers ---
        trace = {}
        oracle.reset()
        model.reset()
        ref = oracle.forward(spec, trace=trace)
        out = model.forward(spec.cuda())
        out = (out[0] if isinstance(out, tuple) else out).cpu()
        # model buffers were rolled after the chunk: interior of conv inputs now holds this chunk's input still
        for name, t in trace.items():
            for b in range(B):
                mine = read(model, name, b)  # [T][F][C]
                want = t[b].permute(2, 1, 0).numpy()  # [C,F,T] -> [T,F,C]
                c = want.shape[2]
                e = rel_err(mine[:, :, :c], want)
                extra = float(np.abs(mine[:, :, c:]).max()) if mine.shape[2] > c else 0.0
                worst = max(worst, e)
                print(f"{name:10s} b={b} rel_err={e:.3e} pad_ch_max={extra:.1e} shape={mine.shape}")
        print("forward rel err vs oracle:", rel_err(out.numpy(), ref.numpy()), " vs golden:",
              rel_err(out.numpy(), g["fwd_chunk1"]))
        # --- end to end ---
        model.reset()
        y = model.realtime_process(torch.from_numpy(mix).cuda())
        y = (y[0] if isinstance(y, tuple) else y).cpu().numpy()
        d = np.abs(y - g["out"]).max()
        print(f"realtime_process max_abs={d:.3e} peak={np.abs(g['out']).max():.3f} "
              f"si_sdr_vs_ref={si_sdr_db(y, g['out']):.1f} dB")
        if "out_cont" in g:
            mix2, _ = synth.make_mixture(B, int(g["meta"][2]) // 2, first_stream=100)
            y2 = model.realtime_process(torch.from_numpy(mix2).cuda(), True).cpu().numpy()
            print(f"continuation max_abs={np.abs(y2 - g['out_cont']).max():.3e}")
        yh = model.realtime_process(torch.from_numpy(mix))
        yh = (yh[0] if isinstance(yh, tuple) else yh).numpy()
        print(f"host-path max_abs vs device path={np.abs(yh - y).max():.3e}")
    return 0 if worst < 1e-1 else 1


if __name__ == "__main__":
    sys.exit(main())
