"""Equivalent of the reference's ``train_distillation.py`` (stage 0, train_distillation.py:126-215) on the B200-native path.

    python -m speech_enhancement_mi_b200.train_distillation DistillationCRN config.yaml [--teacher TemporalCRN.pth]
                                                            [--resume] [--user_defined_name model] [--steps N]

Same flow as the reference: ``model = DistillationCRN(**config['TemporalCRN'], path=<teacher checkpoint>)``
(train_distillation.py:58-59 with ``sub_name = {'DistillationCRN': 'TemporalCRN'}``), Adam over the parameters that
require a gradient, per batch ``loss, stoi, sisnr = model(mixture, source[:, 0], length, flag)`` ->
``(loss / gradient_accumulation).backward()`` -> clip + step every ``gradient_accumulation`` batches
(train_distillation.py:190-200); the dev pass runs the same forward under ``no_grad`` WITHOUT leaving train mode, as the
reference does (train_distillation.py:201-203).  Teacher and student forward / backward are native
(se_crn_train_forward / se_crn_train_tap / se_crn_train_backward_taps); torch only runs the connector loss and the
optimizer.  Checkpoints: ``<dillation_dir or checkpoint_dir>/distillation/<name>/DistillationCRN.pth`` (teacher.*,
student.*, connectors.* -- what predict_distillation.py:33-34 loads) plus optimizer / scheduler / Epoch files.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os

import torch

from . import distillation_crn
from .train import Processor


class DistillProcessor(Processor):
    def __init__(self, args):
        args.engine = "autograd"  # the torch loop of the reference over the native autograd nodes
        super().__init__(args)
        root = self.config["config"].get("dillation_dir", self.config["config"]["checkpoint_dir"])
        self.stage_dir = os.path.join(root, "distillation", args.user_defined_name)

    def build_model(self):
        kw = dict(self.config["TemporalCRN"])
        if self.args.teacher:
            kw["path"] = self.args.teacher
        return distillation_crn.DistillationCRN(**kw)

    def step_loss(self, mixture, source, length, flag):
        with contextlib.redirect_stdout(io.StringIO()):
            return self.model(mixture, source[:, 0], length, flag)

    def eval_mode(self):  # train_distillation.py:201-203 never calls model.eval(): BatchNorm keeps batch statistics
        pass


def main(argv=None):
    ap = argparse.ArgumentParser(description="DistillationCRN training (teacher -> student) on the B200-native path")
    ap.add_argument("name", help="model name: DistillationCRN")
    ap.add_argument("config_path")
    ap.add_argument("--teacher", default=None, help="teacher checkpoint (TemporalCRN.pth); frozen when given")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--user_defined_name", default="model")
    ap.add_argument("--steps", type=int, default=1 << 30)
    ap.add_argument("--items", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=None)
    args = ap.parse_args(argv)
    if args.name != "DistillationCRN":
        raise SystemExit("train_distillation trains DistillationCRN")
    return DistillProcessor(args).train(resume=args.resume)


if __name__ == "__main__":
    main()
