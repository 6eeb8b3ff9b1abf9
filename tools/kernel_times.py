"""GPU: per-kernel device times of one chunk step (CUDA events around each launch, 5 repetitions, L2 flushed between).

    python tools/kernel_times.py [teacher|student] [streams] [fp16|tf32|fp32]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from speech_enhancement_mi_b200._native import check, lib  # noqa: E402
from tools import bench_parts  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "teacher"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
model, _ = bench_parts.build_crn(which, prec, B)
sig = torch.from_numpy(bench_parts.synthetic_signal(B, 4 * 1600)).cuda()
out = torch.empty((B, 1600), device="cuda")
for i in range(3):
    model.process_chunk(sig[:, :, i * 1600 // 2:i * 1600 // 2 + 3200], out)
torch.cuda.synchronize()
L = lib()
total = 0.0
for i in range(L.se_crn_num_kernels(model._ctx)):
    name = C.create_string_buffer(96)
    check(L.se_crn_kernel_info(model._ctx, i, name, 96, None, None, None), "kernel_info")
    ms = C.c_float(0)
    check(L.se_crn_time_kernel(model._ctx, i, B, 5, C.byref(ms)), "time")
    total += ms.value
    print(f"{name.value.decode():40s} {ms.value * 1e3:8.1f} us")
print(f"{'sum':40s} {total * 1e3:8.1f} us")
