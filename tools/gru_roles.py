"""GPU diagnostic: where do the warp roles of the GRU recurrence kernel (gru_wave.cu) wait?

    tools/build_variant.sh prof -DSE_GRU_PROFILE=1 -DSE_GEMM_PROFILE=1
    cp variants/libse_prof.so speech_enhancement_mi_b200/libse_b200.so     # on the GPU box
    python tools/gru_roles.py [teacher|student] [streams]

Per layer (wavefront) or per launch (one-layer form): cycles per step of the MMA warp and the share of each role's time
spent waiting -- producer for the published state (counter) / for a free stage, MMA warp for operands / for the drained
accumulator, epilogue for the accumulator."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from speech_enhancement_mi_b200._native import check, lib  # noqa: E402
from tools import bench_parts  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "teacher"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
model, _ = bench_parts.build_crn(which, "fp16", B)
sig = torch.from_numpy(bench_parts.synthetic_signal(B, 4 * 1600)).cuda()
out = torch.empty((B, 1600), device="cuda")
for i in range(3):
    model.process_chunk(sig[:, :, i * 1600 // 2:i * 1600 // 2 + 3200], out)
torch.cuda.synchronize()
L = lib()
cnt = (C.c_uint64 * 16)()
for i in range(L.se_crn_num_kernels(model._ctx)):
    name = C.create_string_buffer(96)
    check(L.se_crn_kernel_info(model._ctx, i, name, 96, None, None, None), "kernel_info")
    if b"recurrence" not in name.value:
        continue
    check(L.se_debug_gru_counters(cnt, 1), "reset")
    ms = C.c_float(0)
    check(L.se_crn_time_kernel(model._ctx, i, B, 3, C.byref(ms)), "time")
    check(L.se_debug_gru_counters(cnt, 0), "read")
    print(f"{name.value.decode()}: {ms.value * 1e3:.1f} us")
    for layer in range(2):
        c = list(cnt)[8 * layer:8 * layer + 8]
        if c[3] == 0:
            continue
        ctas = max(c[7], 1)  # steps summed over CTAs
        print(f"  layer {layer}: MMA-warp cycles/step {c[3] / ctas:.0f} | MMA wait-operands {c[2] / c[3]:.2f} "
              f"wait-drained-acc {c[6] / c[3]:.2f} | producer wait-state {c[0] / c[3]:.2f} wait-stage {c[1] / c[3]:.2f} "
              f"| epilogue wait-acc {c[4] / max(c[5], 1):.2f}")
