"""Conflict-free equivalent of the reference's ``train.py`` (which does not parse: merge markers at train.py:86-89,
165-169) for the CRN_ELU denoise stage, on the native training path.

    python -m speech_enhancement_mi_b200.train TemporalCRN config.yaml [--resume] [--user_defined_name model]
                                               [--engine native|autograd] [--steps N]

Same flow as train.py:128-260: build ``TemporalCRN(**config['TemporalCRN'])``; Adam(lr, betas=(0.9, 0.999));
ReduceLROnPlateau(factor 0.5, patience 2, min_lr 1e-7) on the dev loss; per batch ``realtime_process`` ->
``compute_loss`` -> ``(loss / gradient_accumulation).backward()`` -> every ``gradient_accumulation`` steps
``clip_grad_norm_(max_grad_norm)`` + ``optimizer.step()``; checkpoints in ``<checkpoint_dir>/denoise/<name>/`` with the
reference's file set (train.py:76-99): ``TemporalCRN.pth`` (state_dict incl. the ``net.0`` alias keys),
``optimizer.pth`` (torch Adam layout), ``scheduler.pth``, ``Epoch.pth``.  Data: ``data_synth.SyntheticPartyDataset``
(the reference corpus is private).  ``--engine native`` (default) drives ``training.NativeTrainer`` (no autograd graph);
``--engine autograd`` runs the reference's literal torch loop on the drop-in model.  Launch with torchrun for data
parallelism (one rank per GPU; gradient all-reduce inside ``NativeTrainer.optimizer_step``).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os

import torch
import yaml

from . import CRN_ELU
from .data_synth import SyntheticPartyDataset
from .training import NativeTrainer


def adam_state_dict(trainer, model, lr):
    """torch.optim.Adam.state_dict() layout (what train.py:90-91 saves) from the flat native moments."""
    offsets = dict(zip([lib_name for lib_name in trainer.param_names], model._t_offsets))
    state, ids = {}, []
    for i, (name, p) in enumerate((n, p) for n, p in model.named_parameters() if p.requires_grad):
        o = offsets[name]
        state[i] = {"step": torch.tensor(float(trainer.step_count)),
                    "exp_avg": trainer.m[o:o + p.numel()].view(p.shape).clone(),
                    "exp_avg_sq": trainer.v[o:o + p.numel()].view(p.shape).clone()}
        ids.append(i)
    group = {"lr": lr, "betas": trainer.betas, "eps": trainer.eps, "weight_decay": 0, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": ids}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(trainer, model, sd):
    offsets = dict(zip(trainer.param_names, model._t_offsets))
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    for i, name in enumerate(names):
        st = sd["state"].get(i)
        if st is None:
            continue
        o = offsets[name]
        n = st["exp_avg"].numel()
        trainer.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
        trainer.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
        trainer.step_count = int(st["step"])
    trainer.lr = float(sd["param_groups"][0]["lr"])


class Processor:
    def __init__(self, args):
        with open(args.config_path, "r", encoding="utf-8") as f:
            self.config = yaml.load(f.read(), Loader=yaml.FullLoader)
        self.args = args
        self.name = args.name
        self.stage_dir = os.path.join(self.config["config"]["checkpoint_dir"], "denoise", args.user_defined_name)
        torch.manual_seed(self.config["config"]["seed"])
        self.model = self.build_model()
        self.epoch, self.train_step, self.dev_step, self.last_loss = -1, 0, 0, 1e8
        dn = self.config["denoise"]
        self.lr, self.accum = float(dn["lr"]), int(dn["gradient_accumulation"])
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if int(os.environ.get("WORLD_SIZE", 1)) > 1 and not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dataset = SyntheticPartyDataset(size=args.items, max_length=self.config["config"]["max_length"],
                                             num_mic=self.config["config"]["num_mic"])
        if args.engine == "native":
            # capacity for the longest piece (pieces with flag=True must find the context of their predecessor), never
            # below what the model's own evaluation path asks for (dev pass of every epoch)
            max_chunks = max(2 * (int(self.config["config"]["max_length"]) // self.model.segment_length + 3),
                             self.model._train_capacity(1, 0))
            self.trainer = NativeTrainer(self.model, lr=self.lr, max_grad_norm=self.config["config"]["max_grad_norm"],
                                         gradient_accumulation=self.accum, device=self.local_rank,
                                         max_chunk_streams=max_chunks)
            self._lr_holder = torch.nn.Parameter(torch.zeros(1))
            self.optimizer = torch.optim.Adam([self._lr_holder], lr=self.lr)  # carries the learning rate for the scheduler
        else:
            self.trainer = None
            self.model.to(f"cuda:{self.local_rank}")
            self.optimizer = torch.optim.Adam((p for p in self.model.parameters() if p.requires_grad), lr=self.lr,
                                              betas=(0.9, 0.999))
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=0.5, patience=2,
                                                                    min_lr=1e-7)

    def build_model(self):
        return getattr(CRN_ELU, self.name)(**self.config[self.name])

    def step_loss(self, mixture, source, length, flag):
        """train.py:195-196 on the drop-in model (autograd engine and every dev pass): (loss, stoi, sisnr) tensors."""
        pred = self.model.realtime_process(mixture, flag)
        with contextlib.redirect_stdout(io.StringIO()):
            return self.model.compute_loss(source[:, 0], pred, length)

    # ---- train.py:76-126 -----------------------------------------------------------------------------------------
    def save_modules(self, with_model):
        if self.rank != 0:
            return
        os.makedirs(self.stage_dir, exist_ok=True)
        if with_model:
            torch.save({k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                       os.path.join(self.stage_dir, self.name + ".pth"))
        opt = adam_state_dict(self.trainer, self.model, self.optimizer.param_groups[0]["lr"]) if self.trainer \
            else self.optimizer.state_dict()
        torch.save(opt, os.path.join(self.stage_dir, "optimizer.pth"))
        torch.save(self.scheduler.state_dict(), os.path.join(self.stage_dir, "scheduler.pth"))
        torch.save({"Epoch": self.epoch, "Train_Step": self.train_step, "Dev_Step": self.dev_step,
                    "Last_Loss": self.last_loss}, os.path.join(self.stage_dir, "Epoch.pth"))

    def load_modules(self):
        self.model.load_state_dict(torch.load(os.path.join(self.stage_dir, self.name + ".pth")), False)
        sd = torch.load(os.path.join(self.stage_dir, "optimizer.pth"))
        if self.trainer:
            self.trainer.reload_parameters()
            load_adam_state_dict(self.trainer, self.model, sd)
            self.optimizer.param_groups[0]["lr"] = self.trainer.lr
        else:
            self.optimizer.load_state_dict(sd)
        self.scheduler.load_state_dict(torch.load(os.path.join(self.stage_dir, "scheduler.pth")))
        p = torch.load(os.path.join(self.stage_dir, "Epoch.pth"))
        self.epoch, self.train_step, self.dev_step, self.last_loss = p["Epoch"], p["Train_Step"], p["Dev_Step"], p["Last_Loss"]

    # ---- train.py:165-236 ----------------------------------------------------------------------------------------
    def run_epoch(self, mode, max_steps):
        self.dataset.set_attribute(mode)
        self.dataset.init_seed(self.epoch + 1 + (10000 if mode == "dev" else 0) + 100 * self.rank)
        dev = f"cuda:{self.local_rank}"
        total, n = 0.0, 0
        for index in range(min(len(self.dataset), max_steps)):
            data = self.dataset[index]
            mixture = data["mix"][None].to(dev)
            source = data["source"][None].to(dev).squeeze(1)
            length = data["length"][None].to(dev).reshape(-1)
            if mode == "train" and self.trainer:
                self.trainer.lr = self.optimizer.param_groups[0]["lr"]
                loss, mae, sisnr = self.trainer.train_step(mixture, source[:, 0].contiguous(), length, data["flag"])
            elif mode == "train":
                self.model.train()
                loss_t, mae, sisnr = self.step_loss(mixture, source, length, data["flag"])
                (loss_t / self.accum).backward()
                if (n + 1) % self.accum == 0:
                    torch.nn.utils.clip_grad_norm_((p for p in self.model.parameters() if p.requires_grad),
                                                   self.config["config"]["max_grad_norm"])
                    self.optimizer.step()
                    self.optimizer.zero_grad()
                loss = float(loss_t)
            else:
                self.eval_mode()
                with torch.no_grad():
                    loss_t, mae, sisnr = self.step_loss(mixture, source, length, data["flag"])
                loss = float(loss_t)
            total += loss
            n += 1
            if mode == "train":
                self.train_step += 1
            else:
                self.dev_step += 1
        return total / max(n, 1)

    def eval_mode(self):
        self.model.eval()

    def train(self, resume=False):
        if resume:
            self.load_modules()
        num_epoch = self.config["denoise"]["num_epoch"] if self.args.epochs is None else self.args.epochs
        for epoch in range(self.epoch + 1, num_epoch):
            train_loss = self.run_epoch("train", self.args.steps)
            dev_loss = self.run_epoch("dev", max(1, self.args.steps // 4))
            # data parallel: every rank evaluated its own shard.  The plateau scheduler and the "best model" decision must
            # see ONE number on every rank, or the ranks halve the learning rate in different epochs and the replicas drift
            # apart (parameters are never re-broadcast after the first step).
            if torch.distributed.is_available() and torch.distributed.is_initialized() \
                    and torch.distributed.get_world_size() > 1:
                t = torch.tensor([dev_loss], dtype=torch.float64, device=f"cuda:{self.local_rank}")
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
                dev_loss = float(t[0]) / torch.distributed.get_world_size()
            improved = dev_loss < self.last_loss
            if improved:
                self.last_loss = dev_loss
            self.scheduler.step(dev_loss)
            self.epoch = epoch
            self.save_modules(with_model=improved)
            if self.rank == 0:
                print(f"epoch {epoch}: train_loss {train_loss:.4f} dev_loss {dev_loss:.4f} "
                      f"lr {self.optimizer.param_groups[0]['lr']:.2e}")
        return self.last_loss


def main(argv=None):
    ap = argparse.ArgumentParser(description="CRN_ELU denoise training on the B200-native path")
    ap.add_argument("name", help="model name: TemporalCRN")
    ap.add_argument("config_path")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--user_defined_name", default="model")
    ap.add_argument("--engine", default="native", choices=["native", "autograd"])
    ap.add_argument("--steps", type=int, default=1 << 30, help="micro-steps per epoch (default: the whole dataset)")
    ap.add_argument("--items", type=int, default=64, help="synthetic dataset size")
    ap.add_argument("--epochs", type=int, default=None)
    args = ap.parse_args(argv)
    return Processor(args).train(resume=args.resume)


if __name__ == "__main__":
    main()
