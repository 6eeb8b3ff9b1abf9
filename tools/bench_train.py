#!/usr/bin/env python
"""CRN_ELU compute_loss training step, data parallel (BASELINE.json configs[4]).

    python tools/bench_train.py [--batch 1] [--seconds 2] [--steps 10] [--precision tf32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

Every rank trains on its own [batch, 3, seconds*16000] synthetic piece: 2 micro-steps (forward, loss, backward) then
one NCCL all-reduce of the flat gradient, clip(5) and Adam(3e-4) -- the sequence of train.py:195-204.  Prints one JSON
line (rank 0): optimizer steps/s, trained audio-seconds per second over all ranks, and the split of the step time."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from speech_enhancement_mi_b200 import CRN_ELU, synth, workload  # noqa: E402
from speech_enhancement_mi_b200.training import NativeTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--model", default="teacher")
ap.add_argument("--graph", action="store_true", help="replay each micro-step as one CUDA graph")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = workload.TEACHER if args.model == "teacher" else dict(num_channels=[8, 8, 16, 16], num_freqs=201, hidden=32,
                                                           num_layers=2, num_inputs=3, kernel_size=3)
model = CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, precision=args.precision, device=local, **cfg)
w = synth.make_crn_weights(seed=0, **cfg)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(w).items()})
B, L = args.batch, int(args.seconds * 16000)
tr = NativeTrainer(model, device=local, gradient_accumulation=2)
mix, src = synth.make_mixture(B, L, first_stream=rank * B)
mix, src = torch.from_numpy(mix).cuda(), torch.from_numpy(src).cuda()
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
t_micro = t_opt = 0.0
losses = []


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for it in range(args.warmup + args.steps):
    if it == args.warmup:
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    ev[0].record()
    for _ in range(2):
        out = tr.micro_step(mix, src, lens, False, check_nan=False, graph=args.graph)
    ev[1].record()
    tr.optimizer_step()
    ev[2].record()
    if it >= args.warmup:
        torch.cuda.synchronize()
        t_micro += ev[0].elapsed_time(ev[1])
        t_opt += ev[1].elapsed_time(ev[2])
        losses.append(0.7 * float(out[0]) - 0.3 * float(out[1]))
e1 = torch.cuda.Event(enable_timing=True)
e1.record()
barrier()
tr.phase_events = []  # one more (untimed) step with per-phase events
for _ in range(2):
    tr.micro_step(mix, src, lens, False, check_nan=False)
tr.optimizer_step()
torch.cuda.synchronize()
phases = {}
for name, a, b in tr.phase_events:
    phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
tr.phase_events = None
ms = e0.elapsed_time(e1) / args.steps
t = torch.tensor([ms], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t)
if rank == 0:
    print(json.dumps({
        "metric": "CRN_ELU training: optimizer steps/s (2 micro-steps of forward+loss+backward, all-reduce, clip, Adam)",
        "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "ms_per_step": ms,
        "trained_audio_s_per_s": world * 2 * B * args.seconds / (ms * 1e-3),
        "ms_micro_steps": t_micro / args.steps, "ms_phases_per_step": phases, "ms_allreduce_clip_adam_rebind": t_opt / args.steps,
        "config": {"model": args.model, "batch_per_rank": B, "piece_seconds": args.seconds, "precision": args.precision,
                   "gradient_accumulation": 2, "cuda_graph": bool(args.graph), "params": int(tr.theta.numel())},
        "loss_first": losses[0], "loss_last": losses[-1], "scaling": "weak", "data": "synthetic"}))
if world > 1:
    dist.destroy_process_group()
