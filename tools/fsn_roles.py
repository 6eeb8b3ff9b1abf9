"""GPU diagnostic: wait-cycle accounting of the TMA tcgen05 GEMM roles over a FullSubNet run (the sub-band LSTM steps).

    tools/build_variant.sh prof -DSE_GEMM_PROFILE=1 && cp variants/libse_prof.so speech_enhancement_mi_b200/libse_b200.so
    python tools/fsn_roles.py [streams]
"""
import ctypes as C
import os
import sys

os.environ["SE_B200_GEMM_PROFILE"] = "1"
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
import torch  # noqa: E402

from speech_enhancement_mi_b200._native import check, lib  # noqa: E402
from tools import bench_parts  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = lib()
cnt = (C.c_uint64 * 8)()
bench_parts.fsn_utterances(B, seconds=0.5, reps=1)  # warm-up incl. graph capture
check(L.se_debug_gemm_counters(cnt, 1), "reset")
r = bench_parts.fsn_utterances(B, seconds=1.0, reps=1)
check(L.se_debug_gemm_counters(cnt, 0), "read")
c = list(cnt)
print(f"FullSubNet, {B} utterances: {r['value']:.0f} audio-s/s, {r['ms_per_chunk_step']:.2f} ms per chunk step")
print(f"TMA GEMM roles: MMA warp wait-operands {c[0] / c[2]:.2f} wait-accumulator {c[1] / c[2]:.2f} | producer wait-stage "
      f"{c[3] / max(c[4], 1):.2f} | epilogue wait-acc {c[5] / max(c[6], 1):.2f} | MMA-warp cycles per tile {c[2] / max(c[7], 1):.0f}")
