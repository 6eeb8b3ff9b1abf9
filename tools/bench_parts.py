"""Measured sub-results that bench.py attaches to its JSON line (and that tools/bench_fsn.py / bench_train.py print alone):

  * ``crn_stream``      the chunk step of the CRN streaming path at any (model, streams, precision): ms per step, p50 / p99
  * ``fsn_utterances``  FullSubNet on 3 s utterances, train=False chunk loop (BASELINE.json configs[3])
  * ``train_step``      CRN_ELU compute_loss training step with the NCCL gradient all-reduce (BASELINE.json configs[4])

Every function times on the device with CUDA events on the current stream after its own warm-up, and takes the max over
ranks when torch.distributed is initialised.  The inputs of every timed loop exceed L2 or change every step.
"""
from __future__ import annotations

import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

RING = 16


def _max_over_ranks(ms, dev):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _barrier():
    import gc
    import torch
    import torch.distributed as dist
    # a model of an earlier sub-benchmark that is only reclaimed by the cycle collector would run its destructor
    # (se_ctx_destroy: a device synchronisation and hundreds of cudaFree) inside the NEXT timed region
    gc.collect()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    torch.cuda.synchronize()


def synthetic_signal(n_streams, n_samples):
    """[n_streams, 3, n_samples] float32 noisy 3-mic streams: 32 base mixtures from synth, tiled with per-stream gains."""
    import numpy as np
    from speech_enhancement_mi_b200 import synth
    base, _ = synth.make_mixture(min(n_streams, 32), n_samples)
    reps = (n_streams + base.shape[0] - 1) // base.shape[0]
    sig = np.tile(base, (reps, 1, 1))[:n_streams].copy()
    gains = 0.5 + 0.5 * synth.uniform01(2021, n_streams, stream=7).astype(np.float32)
    sig *= gains[:, None, None]
    return sig


def build_crn(model_name, precision, max_streams, device=None):
    import torch
    from speech_enhancement_mi_b200 import CRN_ELU, distillation_crn, synth, workload
    cfg = workload.TEACHER if model_name == "teacher" else workload.STUDENT
    cls = CRN_ELU.TemporalCRN if model_name == "teacher" else distillation_crn.TemporalCRN
    kw = {} if device is None else {"device": device}
    model = cls(segment_length=3200, dropout=0.0, precision=precision, max_streams=max_streams, **cfg, **kw)
    w = synth.make_crn_weights(seed=0, **cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(w).items()}, strict=True)
    return model.eval(), cfg


def crn_stream(model_name, streams, precision, steps, warmup=3, latency_steps=0, dev=None):
    """One chunk step for `streams` concurrent streams: device ms per step (max over ranks) and latency percentiles."""
    import torch
    from speech_enhancement_mi_b200 import workload
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    model, cfg = build_crn(model_name, precision, streams, dev.index)
    sig = torch.from_numpy(synthetic_signal(streams, (RING + 1) * 1600)).to(dev)
    out = torch.empty((streams, 1600), dtype=torch.float32, device=dev)

    def view(i):
        off = (i % RING) * 1600
        return sig[:, :, off:off + 3200]

    for i in range(max(warmup, 3)):
        model.process_chunk(view(i), out)
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        model.process_chunk(view(k), out)
    e1.record()
    _barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    res = {"model": model_name, "streams_per_gpu": streams, "precision": precision, "steps": steps, "ms_per_step": ms,
           "value": _world() * streams * workload.AUDIO_SEC_PER_STEP / (ms * 1e-3), "unit": "audio-s/s",
           "algorithmic_tflop_per_s": workload.algorithmic_flops(**cfg)["total"] * streams / (ms * 1e-3) / 1e12}
    if latency_steps:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(latency_steps + 1)]
        ev[0].record()
        for k in range(latency_steps):
            model.process_chunk(view(k), out)
            ev[k + 1].record()
        torch.cuda.synchronize()
        lat = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(latency_steps))
        res.update(p50_chunk_latency_ms=lat[len(lat) // 2], p99_chunk_latency_ms=lat[min(len(lat) - 1, int(0.99 * len(lat)))],
                   latency_steps=latency_steps)
    model._destroy_ctx()  # free the native context now, not when the collector gets to it
    del model
    return res


FSN_CFG = dict(num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15, fb_num_neighbors=0, num_layers=2)
FSN_MFLOP_PER_STREAM_CHUNK = 15547.6  # SURVEY.md section 8(d), hook-counted on the reference


def fsn_utterances(streams, seconds=3.0, reps=2, precision="fp16", peak_tflops=None, dev=None):
    """FullSubNet.realtime_process(train=False) on `streams` utterances of `seconds` (fullsubnet.py:903-961)."""
    import numpy as np
    import torch
    from speech_enhancement_mi_b200 import fullsubnet, synth
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    m = fullsubnet.FullSubNet(num_freqs=201, look_ahead=0, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
                              fb_output_activate_function="ReLU", sb_output_activate_function=False,
                              fb_model_hidden_size=512, sb_model_hidden_size=384, num_mics=3, num_layers=2,
                              weight_init=False, sample_rate=16000, segment_length=3200, win_length=25, hop_length=10,
                              n_fft=400, max_streams=streams, precision=precision, device=dev.index)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_fsn_weights(seed=5, **FSN_CFG).items()})
    B, L = streams, int(seconds * 16000)
    base, _ = synth.make_mixture(min(B, 16), L)
    mix = torch.from_numpy(np.tile(base, ((B + base.shape[0] - 1) // base.shape[0], 1, 1))[:B].copy()).to(dev)
    with torch.cuda.device(dev):
        m.realtime_process(mix, None, flag=False, train=False)  # warm-up
        _barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m.realtime_process(mix, None, flag=False, train=False)
        e1.record()
        _barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1) / reps, dev)
    n_chunks = 2 * (L + 1600 + (3200 - (1600 + (L + 1600) % 3200) % 3200) + 1600) // 3200
    tf = FSN_MFLOP_PER_STREAM_CHUNK * 1e6 * B * n_chunks / (ms * 1e-3) / 1e12
    res = {"metric": "enhanced audio-sec/sec (FullSubNet, chunked train=False path)",
           "value": _world() * B * seconds / (ms * 1e-3), "unit": "audio-s/s", "streams_per_gpu": B, "utterance_s": seconds,
           "chunks": n_chunks, "ms_per_utterance_batch": ms, "ms_per_chunk_step": ms / n_chunks,
           "algorithmic_tflop_per_s_per_gpu": tf, "precision": precision}
    if peak_tflops:
        res["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peak_tflops, "unit": "TFLOP/s", "frac": tf / peak_tflops}
    m._destroy()
    del m
    return res


def train_step(batch=1, seconds=2.0, steps=6, warmup=2, precision="tf32", graph=True, dev=None):
    """train.py:195-204 per rank on its own piece: 2 micro-steps (forward, loss, backward), ONE summing all-reduce of the
    flat gradient over the ranks (NCCL when launched under torchrun), clip(5), Adam(3e-4)."""
    import torch
    from speech_enhancement_mi_b200 import CRN_ELU, synth, workload
    from speech_enhancement_mi_b200.training import NativeTrainer
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    rank = int(os.environ.get("RANK", 0))
    cfg = workload.TEACHER
    model = CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, precision=precision, device=dev.index, **cfg)
    w = synth.make_crn_weights(seed=0, **cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(w).items()})
    B, L = batch, int(seconds * 16000)
    with torch.cuda.device(dev):
        tr = NativeTrainer(model, device=dev.index, gradient_accumulation=2)
        mix, src = synth.make_mixture(B, L, first_stream=rank * B)
        mix, src = torch.from_numpy(mix).to(dev), torch.from_numpy(src).to(dev)
        lens = torch.full((B,), L, dtype=torch.int32, device=dev)
        # no host synchronisation inside the timed region (the K steps are bracketed by a barrier + synchronize on both
        # sides): a per-step synchronize left the GPU idle while the host -- and, data parallel, the slower rank's host --
        # prepared the next step (2 ranks: 10.8 ms per step against 7.7 ms of device time)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0 = None
        for it in range(warmup + steps):
            if it == warmup:
                _barrier()
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            for _ in range(2):
                tr.micro_step(mix, src, lens, False, check_nan=False, graph=graph)
            if it >= warmup:
                ev[it - warmup][0].record()
            tr.optimizer_step()
            if it >= warmup:
                ev[it - warmup][1].record()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _barrier()
        t_opt = sum(a.elapsed_time(b) for a, b in ev)
        ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev)
        n_theta = int(tr.theta.numel())
    res = {"metric": "CRN_ELU training: optimizer steps/s (2 micro-steps of forward+loss+backward, all-reduce, clip, Adam)",
           "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms,
           "trained_audio_s_per_s": _world() * 2 * B * seconds / (ms * 1e-3),
           "ms_allreduce_clip_adam_rebind": t_opt / steps, "allreduce_bytes": 4 * n_theta,
           "collective": "NCCL all-reduce(sum) of the flat fp32 gradient" if _world() > 1 else "none (one rank)",
           "config": {"batch_per_rank": B, "piece_seconds": seconds, "precision": precision, "gradient_accumulation": 2,
                      "cuda_graph": bool(graph), "params": n_theta}}
    del tr
    model._destroy_ctx()
    del model
    return res
