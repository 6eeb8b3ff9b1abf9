"""GPU diagnostic: where do the warp roles of the TMA tcgen05 GEMMs wait?  (SE_B200_GEMM_PROFILE=1 cycle counters)

    python tools/gemm_roles.py [streams]

For every GEMM kernel of the chunk step: device time, and the share of the MMA thread's time spent waiting for operand
stages / for a drained accumulator, of the producer's time waiting for a free stage, of the epilogue's time waiting for an
accumulator.  Counters are per role, summed over CTAs."""
import ctypes as C
import os
import sys

os.environ["SE_B200_GEMM_PROFILE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from speech_enhancement_mi_b200._native import check, lib  # noqa: E402
from tools import bench_parts  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model, _ = bench_parts.build_crn("teacher", "fp16", B)
sig = torch.from_numpy(bench_parts.synthetic_signal(B, 4 * 1600)).cuda()
out = torch.empty((B, 1600), device="cuda")
for i in range(3):
    model.process_chunk(sig[:, :, i * 1600 // 2:i * 1600 // 2 + 3200], out)
torch.cuda.synchronize()
L = lib()
cnt = (C.c_uint64 * 8)()
for i in range(L.se_crn_num_kernels(model._ctx)):
    name = C.create_string_buffer(96)
    check(L.se_crn_kernel_info(model._ctx, i, name, 96, None, None, None), "kernel_info")
    check(L.se_debug_gemm_counters(cnt, 1), "reset")
    ms = C.c_float(0)
    check(L.se_crn_time_kernel(model._ctx, i, B, 3, C.byref(ms)), "time")
    check(L.se_debug_gemm_counters(cnt, 0), "read")
    c = list(cnt)
    if c[2] == 0:
        continue
    print(f"{name.value.decode():28s} {ms.value * 1e3:7.1f} us  MMA: wait-operands {c[0] / c[2]:.2f} wait-accumulator {c[1] / c[2]:.2f} "
          f"| producer wait-stage {c[3] / max(c[4], 1):.2f} | epilogue wait-acc {c[5] / max(c[6], 1):.2f}  "
          f"cycles/tile (MMA thread) {c[2] / max(c[7], 1):.0f}")
