#!/usr/bin/env python
"""Offline enhancement of a few long signals (predict.py:92 use case): serial chunk loop vs the chunk-batched forward.

    python tools/bench_file.py [--streams 1] [--seconds 10] [--precision tf32]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from speech_enhancement_mi_b200 import CRN_ELU, synth, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
model = CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, precision=args.precision, **workload.TEACHER)
w = synth.make_crn_weights(seed=0, **workload.TEACHER)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(w).items()})
model.eval().cuda()
L = int(args.seconds * 16000)
mix, _ = synth.make_mixture(args.streams, L)
x = torch.from_numpy(mix).cuda()
res = {}
for mode in (False, True):
    model.chunk_batch = mode
    y = model.realtime_process(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        y = model.realtime_process(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    res["chunk_batched" if mode else "serial_chunk_loop"] = {"ms": ms, "x_realtime": args.streams * args.seconds / (ms * 1e-3)}
    res["out_" + str(mode)] = y
err = float((res.pop("out_True") - res.pop("out_False")).abs().max())
print(json.dumps({"metric": "offline realtime_process of long signals, 1 GPU", "streams": args.streams,
                  "seconds": args.seconds, "precision": args.precision, "max_abs_difference": err, **res}))
