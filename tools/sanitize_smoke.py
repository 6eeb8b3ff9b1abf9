"""Small invocation of the hot path for compute-sanitizer (tools/sanitize.sh): the chunk loop of `realtime_process` on two
streams x three chunks in the precision given on the command line (default fp16 = every round-2 kernel: fused
pre-convolutions, mma.sync encoder / decoder blocks, TMA GEMMs, persistent GRU), plus the flag=True continuation.
No oracle, no timing: the sanitizer slows kernels 10-100x."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from speech_enhancement_mi_b200 import CRN_ELU, synth  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
tag = sys.argv[2] if len(sys.argv) > 2 else "small"
cfg = dict(num_channels=[8, 8, 16, 16], num_freqs=201, hidden=32, num_layers=2, num_inputs=3, kernel_size=3) \
    if tag == "small" else dict(num_channels=[16, 32, 64, 128], num_freqs=201, hidden=512, num_layers=2, num_inputs=3,
                                kernel_size=3)
model = CRN_ELU.TemporalCRN(segment_length=3200, dropout=0.0, precision=precision, **cfg)
w = synth.make_crn_weights(seed=7, **cfg)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.with_alias_keys(w).items()}, strict=True)
model.eval()
model.chunk_batch = False  # the streaming chunk loop (the path the bench times)
mix, _ = synth.make_mixture(2, 3200)
with torch.no_grad():
    y = model.realtime_process(torch.from_numpy(mix).cuda())
    y2 = model.realtime_process(torch.from_numpy(mix[:, :, :1600].copy()).cuda(), True)
torch.cuda.synchronize()
print("sanitize_smoke", precision, tag, tuple(y.shape), tuple(y2.shape), float(y.abs().max()), bool(torch.isfinite(y).all()))
