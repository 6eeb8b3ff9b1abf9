"""Training step of the reference's train loop (train.py:195-204; config.yaml:10-11,92,99-100) on the native path:

    pred = model.realtime_process(mixture, flag); loss, mae, sisnr = model.compute_loss(source[:, 0], pred, length)
    (loss / gradient_accumulation).backward()
    every gradient_accumulation micro-steps: clip_grad_norm_(5) -> Adam(lr=3e-4).step() -> zero_grad()

``NativeTrainer`` runs exactly that sequence through the C-ABI (se_crn_train_forward, se_loss_terms_grad,
se_crn_train_backward, se_clip_adam_step, se_crn_bind_weights_flat) on ONE flat parameter / gradient vector, without
building an autograd graph.  Data parallelism (BASELINE.json configs[4]): every rank owns its own utterance pieces;
the only exchange is one summing all-reduce of the flat gradient (``torch.distributed``, NCCL on GPUs) per optimizer
step, followed by the identical clip + Adam on every rank.

PyTorch is used for device memory, streams and the collective only.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native
from ._native import check, lib


def flat_layout(names, numels):
    """Offsets of each tensor in the flat vector (registry order) -- pure host logic, shared with the tests."""
    offs, o = [], 0
    for n in numels:
        offs.append(o)
        o += int(n)
    return dict(zip(names, offs)), o


def allreduce_mean_(flat: torch.Tensor, group=None):
    """Sum the flat gradient over the data-parallel ranks; returns the scale (1 / world) the optimizer applies.
    One collective per optimizer step (24.6 MB for the teacher)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class NativeTrainer:
    def __init__(self, model, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=5.0, gradient_accumulation=2,
                 device=None, max_chunk_streams=None, group=None):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = float(max_grad_norm)
        self.accum = int(gradient_accumulation)
        self.group = group
        self.device = int(device if device is not None else model._pick_device(None))
        self.micro = 0
        self.step_count = 0
        self._graphs = {}
        self.phase_events = None  # set to [] to collect (name, start, end) CUDA events of the micro-step phases
        if max_chunk_streams:
            model._max_streams = max(model._max_streams, int(max_chunk_streams))
        with torch.cuda.device(self.device):
            model.to(f"cuda:{self.device}")
            # sized like the model's own evaluation path sizes it (44 chunk-streams per utterance), so that the dev pass of
            # an epoch (model.realtime_process -> chunk-batched forward) never has to re-create the context
            model._ensure_train_ctx(model._train_capacity(1, 0), self.device, keep_state=False)
            self._ctx_gen = model._tctx_gen
            params = model._train_params()
            self.param_names = [lib().se_crn_param_name(self.ctx, i).decode() for i in range(len(params))]
            n = lib().se_crn_num_theta(self.ctx)
            dev = torch.device("cuda", self.device)
            self.theta = torch.cat([p.detach().to(device=dev, dtype=torch.float32).reshape(-1) for p in params])
            assert self.theta.numel() == n
            # the module's parameters become views of the flat vector: state_dict() always shows the trained weights
            for p, off in zip(params, model._t_offsets):
                p.data = self.theta[off:off + p.numel()].view(p.shape)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
            self.gtmp = torch.empty(n, dtype=torch.float32, device=dev)
            self.m = torch.zeros(n, dtype=torch.float32, device=dev)
            self.v = torch.zeros(n, dtype=torch.float32, device=dev)
            self.one = torch.ones(1, dtype=torch.float32, device=dev)
            self.w_stoi = torch.full((1,), 0.7 / self.accum, dtype=torch.float32, device=dev)    # CRN_ELU.py:529
            self.w_sisnr = torch.full((1,), -0.3 / self.accum, dtype=torch.float32, device=dev)  # sisnr enters negated
            self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
            self._rebind()

    @property
    def ctx(self):
        """The model's CURRENT native training context.  The handle is never cached: the model re-creates the context
        when a call needs more chunk-streams (the old handle is then freed); CUDA graphs captured on the old context and
        its weight binding die with it."""
        if self.model._tctx is None:
            self.model._ensure_train_ctx(self.model._train_capacity(1, 0), self.device, keep_state=False)
        if self.model._tctx_gen != self._ctx_gen:
            self._ctx_gen = self.model._tctx_gen
            self._graphs.clear()
            with torch.cuda.device(self.device):
                self._rebind()
        return self.model._tctx

    def reload_parameters(self):
        """After model.load_state_dict (resume, train.py:110): copy_ wrote through the views into the flat vector; re-lay
        the weights out."""
        with torch.cuda.device(self.device):
            self._rebind()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _mark(self, name=None, start=None):
        if self.phase_events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if name is not None:
            self.phase_events.append((name, start, e))
        return e

    def _rebind(self):
        check(lib().se_crn_bind_weights_flat(self.model._tctx, self.theta.data_ptr(), self._stream()),
              "se_crn_bind_weights_flat")
        self.model._tbound_versions = tuple((t.data_ptr(), t._version) for t in self.model._train_params())

    def _launch_micro(self, mixture, source, lens, B, L, flag, pred, out2, d_stoi, d_sisnr):
        """The native calls of one micro-step, all on the current stream (capturable into a CUDA graph)."""
        st = self._stream()
        mark = self._mark
        t0 = mark()
        check(lib().se_crn_train_forward(self.ctx, mixture.data_ptr(), B, L, int(bool(flag)), pred.data_ptr(), st),
              "se_crn_train_forward")
        self.model._fwd_gen += 1
        t1 = mark("forward", t0)
        check(lib().se_loss_terms_grad(source.data_ptr(), pred.data_ptr(), lens.data_ptr(), B, L, out2.data_ptr(),
                                       d_stoi.data_ptr(), d_sisnr.data_ptr(), st), "se_loss_terms_grad")
        t2 = mark("loss+grad", t1)
        check(lib().se_axpby_dev(self.w_stoi.data_ptr(), d_stoi.data_ptr(), self.w_sisnr.data_ptr(), d_sisnr.data_ptr(),
                                 d_stoi.data_ptr(), d_stoi.numel(), st), "se_axpby_dev")
        check(lib().se_crn_train_backward(self.ctx, d_stoi.data_ptr(), self.gtmp.data_ptr(), st), "se_crn_train_backward")
        check(lib().se_axpby_dev(self.one.data_ptr(), self.grad.data_ptr(), self.one.data_ptr(), self.gtmp.data_ptr(),
                                 self.grad.data_ptr(), self.grad.numel(), st), "se_axpby_dev")
        mark("backward", t2)

    def micro_step(self, mixture, source, length, flag=False, check_nan=True, graph=False):
        """One forward + backward; gradients accumulate.  Returns a device tensor [stoi_loss, sisnr_db].

        ``graph=True`` replays the micro-step as ONE CUDA graph (captured on first use per (B, L, flag); inputs are
        staged into fixed device buffers).  A NaN loss is then not filtered (check_nan needs the eager path)."""
        B, _, L = mixture.shape
        with torch.cuda.device(self.device):
            _, n_chunks = _native.chunk_grid(L + (0 if flag else self.model.segment_length // 2), self.model.segment_length)
            if self.model._tctx is None or B * n_chunks > self.model._tctx_capacity:
                self.model._ensure_train_ctx(self.model._train_capacity(B, n_chunks), self.device, keep_state=bool(flag))
            _ = self.ctx  # notices a re-created context (here or by the model's evaluation path): graphs dropped, weights re-bound
            dev = self.theta.device
            lens = torch.as_tensor(length)
            if lens.device != dev or lens.dtype != torch.int32:
                lens = lens.to(device=dev, dtype=torch.int32)
            if graph:
                key = (B, L, bool(flag))
                g = self._graphs.get(key)
                if g is None:
                    bufs = dict(mix=torch.empty_like(mixture), src=torch.empty_like(source), lens=torch.empty_like(lens),
                                pred=torch.empty((B, L), dtype=torch.float32, device=dev),
                                out2=torch.empty(2, dtype=torch.float32, device=dev),
                                d_stoi=torch.empty((B, L), dtype=torch.float32, device=dev),
                                d_sisnr=torch.empty((B, L), dtype=torch.float32, device=dev))
                    for k, v in (("mix", mixture), ("src", source), ("lens", lens)):
                        bufs[k].copy_(v)
                    if not flag:  # eager warm-up (one-time lazy setup inside the library); flag=False resets the state
                        saved = self.grad.clone()
                        self._launch_micro(bufs["mix"], bufs["src"], bufs["lens"], B, L, flag, bufs["pred"], bufs["out2"],
                                           bufs["d_stoi"], bufs["d_sisnr"])
                        self.grad.copy_(saved)
                    torch.cuda.synchronize()
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        self._launch_micro(bufs["mix"], bufs["src"], bufs["lens"], B, L, flag, bufs["pred"], bufs["out2"],
                                           bufs["d_stoi"], bufs["d_sisnr"])
                    g = self._graphs[key] = (cg, bufs)
                cg, bufs = g
                bufs["mix"].copy_(mixture, non_blocking=True)
                bufs["src"].copy_(source, non_blocking=True)
                bufs["lens"].copy_(lens, non_blocking=True)
                cg.replay()
                self.micro += 1
                self.last_pred = bufs["pred"]
                return bufs["out2"]
            pred = torch.empty((B, L), dtype=torch.float32, device=dev)
            out2 = torch.empty(2, dtype=torch.float32, device=dev)
            d_stoi = torch.empty((B, L), dtype=torch.float32, device=dev)
            d_sisnr = torch.empty((B, L), dtype=torch.float32, device=dev)
            if check_nan:  # CRN_ELU.py:531-534: a NaN loss is zero-filled and contributes no gradient
                st = self._stream()
                check(lib().se_crn_train_forward(self.ctx, mixture.data_ptr(), B, L, int(bool(flag)), pred.data_ptr(), st),
                      "se_crn_train_forward")
                self.model._fwd_gen += 1
                check(lib().se_loss_terms_grad(source.data_ptr(), pred.data_ptr(), lens.data_ptr(), B, L, out2.data_ptr(),
                                               d_stoi.data_ptr(), d_sisnr.data_ptr(), st), "se_loss_terms_grad")
                self.micro += 1
                self.last_pred = pred
                if bool(torch.isnan(out2).any()):
                    return torch.zeros_like(out2)
                check(lib().se_axpby_dev(self.w_stoi.data_ptr(), d_stoi.data_ptr(), self.w_sisnr.data_ptr(),
                                         d_sisnr.data_ptr(), d_stoi.data_ptr(), d_stoi.numel(), st), "se_axpby_dev")
                check(lib().se_crn_train_backward(self.ctx, d_stoi.data_ptr(), self.gtmp.data_ptr(), st),
                      "se_crn_train_backward")
                check(lib().se_axpby_dev(self.one.data_ptr(), self.grad.data_ptr(), self.one.data_ptr(),
                                         self.gtmp.data_ptr(), self.grad.data_ptr(), self.grad.numel(), st), "se_axpby_dev")
                return out2
            self._launch_micro(mixture, source, lens, B, L, flag, pred, out2, d_stoi, d_sisnr)
        self.micro += 1
        self.last_pred = pred
        return out2

    def optimizer_step(self):
        """All-reduce (data parallel), clip_grad_norm_(max_grad_norm), Adam, re-layout of the weights, zero_grad."""
        with torch.cuda.device(self.device):
            _ = self.ctx  # re-bind first if the evaluation path re-created the context since the last micro-step
            scale = allreduce_mean_(self.grad, self.group)
            self.step_count += 1
            check(lib().se_clip_adam_step(self.theta.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                          self.theta.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                          self.step_count, self.max_grad_norm, scale, self.norm.data_ptr(), self._stream()),
                  "se_clip_adam_step")
            self._rebind()
            self.grad.zero_()
        return self.norm

    def train_step(self, mixture, source, length, flag=False):
        """train.py:195-204 for one batch: micro-step, and every `gradient_accumulation`-th call the optimizer step."""
        out = self.micro_step(mixture, source, length, flag)
        if self.micro % self.accum == 0:
            self.optimizer_step()
        loss = 0.7 * float(out[0]) - 0.3 * float(out[1])
        return (loss if not math.isnan(loss) else 0.0), float(out[0]), -float(out[1])
