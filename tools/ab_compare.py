"""GPU diagnostic (test infrastructure): the same model with an A/B switch of the library on and off.

    python tools/ab_compare.py SE_B200_FRONT_MMA,SE_B200_ENC_MMA [tag ...] [--precision fp16] [--time]

For each tag (default: crn_small crn_student crn_teacher) builds the model twice -- all named switches = 0 (the older
kernels) and switches unset (the default path) -- runs `realtime_process` (+ the flag=True continuation) on the fixture
input and prints: new-vs-old max-abs, each arm's max-abs / SI-SDR against the reference fixture, and per-layer tensors of
the one-chunk forward where both arms expose them.  Exit code 1 when the new arm is not within the stated fp16 tolerance.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from common import TOL, load_golden, make_model, si_sdr_db  # noqa: E402
from oracle import synth  # noqa: E402
from speech_enhancement_mi_b200._native import lib  # noqa: E402


def read(model, name, b):
    buf = np.zeros(4 * 1024 * 1024, dtype=np.float32)
    dims = (C.c_int * 3)()
    if lib().se_debug_read(model._ctx, name.encode(), b, buf.ctypes.data, buf.size, dims) != 0:
        return None
    t, f, c = dims[0], dims[1], dims[2]
    return buf[: t * f * c].reshape(t, f, c).copy()


def run(tag, precision, switches, off):
    for s in switches:
        if off:
            os.environ[s] = "0"
        else:
            os.environ.pop(s, None)
    g = load_golden(tag)
    model = make_model(tag, precision)
    B, L = int(g["meta"][1]), int(g["meta"][2])
    mix, _ = synth.make_mixture(B, L)
    res = {}
    with torch.no_grad():
        spec = torch.from_numpy(g["spec_chunk1"]).cuda()
        model.reset()
        f = model.forward(spec)
        f = (f[0] if isinstance(f, tuple) else f).cpu().numpy()
        res["fwd"] = f
        for name in ["pre_in0", "enc_in0", "enc_in1", "enc_in2", "enc_in3", "xg", "dec_in0", "dec_in1", "dec_in2", "dec_in3",
                     "ylast"]:
            t = read(model, name, 0)
            if t is not None:
                res["L:" + name] = t
        model.reset()
        y = model.realtime_process(torch.from_numpy(mix).cuda())
        res["out"] = (y[0] if isinstance(y, tuple) else y).cpu().numpy()
        if "out_cont" in g:
            mix2, _ = synth.make_mixture(B, L // 2, first_stream=100)
            y2 = model.realtime_process(torch.from_numpy(mix2).cuda(), True)
            res["out_cont"] = (y2[0] if isinstance(y2, tuple) else y2).cpu().numpy()
    del model
    return res, g


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    precision = "fp16"
    if "--precision" in sys.argv:
        precision = sys.argv[sys.argv.index("--precision") + 1]
        args = [a for a in args if a != precision]
    switches = args[0].split(",")
    tags = args[1:] or ["crn_small", "crn_student", "crn_teacher"]
    bad = False
    for tag in tags:
        old, g = run(tag, precision, switches, off=True)
        new, _ = run(tag, precision, switches, off=False)
        peak = float(np.abs(g["out"]).max())
        print(f"== {tag} [{precision}] switches {switches}: peak {peak:.3f}")
        for k in sorted(new):
            if k.startswith("L:") and k in old:
                a, b = new[k], old[k]
                print(f"   {k[2:]:8s} new-vs-old max_abs={np.abs(a - b).max():.3e}  (tensor peak {np.abs(b).max():.3e})"
                      f"{'  NaN!' if not np.isfinite(a).all() else ''}")
        for key in ("fwd", "out", "out_cont"):
            if key not in new:
                continue
            ref = g["fwd_chunk1"] if key == "fwd" else g[key]
            pk = float(np.abs(ref).max())
            dn, do = float(np.abs(new[key] - ref).max()), float(np.abs(old[key] - ref).max())
            line = f"   {key:8s} new-vs-old={np.abs(new[key] - old[key]).max():.3e}  new-vs-ref={dn:.3e} old-vs-ref={do:.3e} (peak {pk:.3f})"
            if key != "fwd":
                line += f"  si_sdr new {si_sdr_db(new[key], ref):.1f} dB old {si_sdr_db(old[key], ref):.1f} dB"
                tol = TOL[precision]
                if not (dn <= tol["wave_max_abs"] * max(1.0, pk) and si_sdr_db(new[key], ref) >= tol["si_sdr_vs_ref_db"]):
                    bad = True
                    line += "  <-- OUT OF TOLERANCE"
            print(line)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
