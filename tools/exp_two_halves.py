"""Experiment: one 1024-stream context against two 512-stream contexts stepped concurrently on two CUDA streams
(same total work per step).  Prints ms per step of both arrangements."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools import bench_parts  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = 50
dev = torch.device("cuda", 0)
RING = bench_parts.RING
sig = torch.from_numpy(bench_parts.synthetic_signal(B, (RING + 1) * 1600)).to(dev)


def view(i, lo, hi):
    off = (i % RING) * 1600
    return sig[lo:hi, :, off:off + 3200]


def run(nparts):
    per = B // nparts
    models = [bench_parts.build_crn("teacher", "fp16", per, 0)[0] for _ in range(nparts)]
    outs = [torch.empty((per, 1600), dtype=torch.float32, device=dev) for _ in range(nparts)]
    streams = [torch.cuda.Stream(dev) for _ in range(nparts)]
    main = torch.cuda.current_stream(dev)

    def step(i):
        for k in range(nparts):
            streams[k].wait_stream(main)
            with torch.cuda.stream(streams[k]):
                models[k].process_chunk(view(i, k * per, (k + 1) * per), outs[k])
        for k in range(nparts):
            main.wait_stream(streams[k])

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{nparts} x {per} streams: {ms:.3f} ms per step = {B * 0.1 / (ms * 1e-3):.0f} audio-s/s", flush=True)
    del models


run(1)
run(parts)
if parts != 2:
    run(2)
