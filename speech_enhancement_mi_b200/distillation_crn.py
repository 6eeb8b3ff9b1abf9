"""Drop-in for the student forward path of the reference's ``distillation_crn.TemporalCRN`` (distillation_crn.py:283-501).

Same graph as CRN_ELU.TemporalCRN with two numerics differences the kernels switch on (SE_VARIANT_DISTILLED):
GlobalLayerNorm divides by ``sqrt(var) + 1e-8`` (distillation_crn.py:51) and the phase is
``arctan(im / (re + 1e-8) + 1e-8)`` (distillation_crn.py:340).  ``forward`` / ``realtime_process`` return a tuple as in
the reference; the second element (intermediate features for the distillation loss, distillation_crn.py:343-377) is a
training-only output and is returned as an empty list -- the distillation trainer is out of scope.
"""
from __future__ import annotations

from . import _native
from .CRN_ELU import TemporalCRN as _TemporalCRN


class TemporalCRN(_TemporalCRN):
    _variant = _native.SE_VARIANT_DISTILLED

    def forward(self, x):
        return super().forward(x), []

    def realtime_process(self, mixture, flag=False):
        return super().realtime_process(mixture, flag), []

    def get_channel_num(self):  # distillation_crn.py:384-385
        c = self.num_channels
        return [c[-1], c[-1], c[2], c[1], c[0]]
