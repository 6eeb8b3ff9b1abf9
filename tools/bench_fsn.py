#!/usr/bin/env python
"""FullSubNet streaming throughput (BASELINE.json configs[3]: 3 s synthetic utterances, train=False chunk loop).

    python tools/bench_fsn.py [--streams B] [--seconds 3] [--reps 3]

Prints one JSON line: enhanced audio-s/s for B concurrent utterances on one GPU, ms per chunk step, and the achieved
tensor throughput of the chunk step against the algorithmic 15,547.6 MFLOP per stream-chunk."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from speech_enhancement_mi_b200 import fullsubnet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=64)
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="fp16", choices=["tf32", "fp16"])
args = ap.parse_args()
cfg = dict(num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15, fb_num_neighbors=0, num_layers=2)
m = fullsubnet.FullSubNet(num_freqs=201, look_ahead=0, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
                          fb_output_activate_function="ReLU", sb_output_activate_function=False,
                          fb_model_hidden_size=512, sb_model_hidden_size=384, num_mics=3, num_layers=2, weight_init=False,
                          sample_rate=16000, segment_length=3200, win_length=25, hop_length=10, n_fft=400,
                          max_streams=args.streams, precision=args.precision)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_fsn_weights(seed=5, **cfg).items()})
B, L = args.streams, int(args.seconds * 16000)
base, _ = synth.make_mixture(min(B, 16), L)
mix = torch.from_numpy(np.tile(base, ((B + base.shape[0] - 1) // base.shape[0], 1, 1))[:B].copy()).cuda()
m.realtime_process(mix, None, flag=False, train=False)  # warm-up (graph capture, weights)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    y = m.realtime_process(mix, None, flag=False, train=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
n_chunks = 2 * (L + 1600 + (3200 - (1600 + (L + 1600) % 3200) % 3200) + 1600) // 3200
flops = 15547.6e6 * B * n_chunks
print(json.dumps({"metric": "enhanced audio-sec/sec (FullSubNet, chunked train=False path)", "value": B * args.seconds / (ms * 1e-3),
                  "unit": "audio-s/s", "streams": B, "utterance_s": args.seconds, "chunks": n_chunks,
                  "ms_per_utterance_batch": ms, "ms_per_chunk_step": ms / n_chunks,
                  "achieved_tflops": flops / (ms * 1e-3) / 1e12, "dtype": args.precision, "data": "synthetic"}))
