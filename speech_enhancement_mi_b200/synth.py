"""Deterministic synthetic weights and 16 kHz multi-mic mixtures.

Synthetic DATA only (shared by tests/, bench.py and oracle/make_golden.py): no arithmetic of the enhancement path lives here.

Everything is generated from a counter-based integer hash (splitmix64) so that the build container, the GPU box and
any later machine produce bit-identical float32 tensors without depending on a framework RNG stream.  Mirrors what the
reference feeds the path: 3-mic mixtures in [-0.95, 0.95] (reference data_c.py:16,249-250), SNR drawn from [-5, 25] dB
(reference config.yaml:52-53), seed = 2021 + stream (reference utility.py:148-151), and parameter tensors with the key
set / shapes of ``TemporalCRN.state_dict()`` (reference CRN_ELU.py:321-365).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n float64 uniforms in [0,1): hash of (seed, stream, counter); exact integer arithmetic."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        base = _splitmix64(np.array([base ^ np.uint64(stream * 0x632BE59BD9B4E019 & 0xFFFFFFFFFFFFFFFF)],
                                    dtype=np.uint64))[0]
        ctr = np.arange(n, dtype=np.uint64)
        bits = _splitmix64(base + ctr * np.uint64(0xD1342543DE82EF95))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def normal(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n float64 standard normals by the sum of 12 uniforms minus 6 (no transcendental => bit-stable everywhere)."""
    u = uniform01(seed, 12 * n, stream).reshape(12, n)
    return u.sum(axis=0) - 6.0


# ----------------------------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------------------------
def crn_param_shapes(num_channels, num_freqs=201, hidden=512, num_layers=2, num_inputs=3, kernel_size=3):
    """Ordered {key: shape} of the DISTINCT parameter tensors of TemporalCRN (reference CRN_ELU.py:335-365).

    The reference state_dict additionally repeats ``*.conv.*`` under the alias ``*.net.0.*`` (CRN_ELU.py:225,282).
    """
    cin0 = 2 * num_inputs - 1
    shapes = OrderedDict()

    def conv_block(prefix, ci, co, kf, kt):
        shapes[f"{prefix}.conv.weight"] = (co, ci, kf, kt)
        shapes[f"{prefix}.conv.bias"] = (co,)
        shapes[f"{prefix}.conv_trans.weight"] = (co, co, 1, 1)
        shapes[f"{prefix}.conv_trans.bias"] = (co,)
        shapes[f"{prefix}.conv_gated.weight"] = (co, co, 1, 1)
        shapes[f"{prefix}.conv_gated.bias"] = (co,)
        shapes[f"{prefix}.norm.weight"] = (1, co, 1, 1)
        shapes[f"{prefix}.norm.bias"] = (1, co, 1, 1)

    for i in range(3):
        conv_block(f"preconvlist.{i}", cin0, cin0, 5, 5)
    L = len(num_channels)
    for i in range(L):
        ci = cin0 if i == 0 else num_channels[i - 1]
        conv_block(f"convlist.{i}", ci, num_channels[i], 5, kernel_size)
    # decoder in execution order: deconvlist.j maps channels[L-1-j] -> channels[L-2-j] (last -> 2)
    for j in range(L):
        ci = num_channels[L - 1 - j]
        co = num_channels[L - 2 - j] if j < L - 1 else 2
        p = f"deconvlist.{j}"
        shapes[f"{p}.conv.weight"] = (ci, co, 5, kernel_size)
        shapes[f"{p}.conv.bias"] = (co,)
        shapes[f"{p}.residualmask.weight"] = (co, co, 1, 1)
        shapes[f"{p}.residualmask.bias"] = (co,)
        shapes[f"{p}.residualnorm.weight"] = (1, co, 1, 1)
        shapes[f"{p}.residualnorm.bias"] = (1, co, 1, 1)
        shapes[f"{p}.residual.weight"] = (co, co, 1, 1)
        shapes[f"{p}.residual.bias"] = (co,)
        shapes[f"{p}.norm.weight"] = (1, co, 1, 1)
        shapes[f"{p}.norm.bias"] = (1, co, 1, 1)
    feat = (num_freqs // 2 ** L + 1) * num_channels[-1]
    for l in range(num_layers):
        isz = feat if l == 0 else hidden
        shapes[f"gru.sequence_model.weight_ih_l{l}"] = (3 * hidden, isz)
        shapes[f"gru.sequence_model.weight_hh_l{l}"] = (3 * hidden, hidden)
        shapes[f"gru.sequence_model.bias_ih_l{l}"] = (3 * hidden,)
        shapes[f"gru.sequence_model.bias_hh_l{l}"] = (3 * hidden,)
    shapes["gru.fc_output_layer.weight"] = (feat, hidden)
    shapes["gru.fc_output_layer.bias"] = (feat,)
    shapes["gru.norm.weight"] = (1, 1, 1, feat)
    shapes["gru.norm.bias"] = (1, 1, 1, feat)
    return shapes


def _fan_in(key: str, shape) -> int:
    if key.endswith("bias") or "norm" in key:
        return 0
    if "deconvlist" in key and ".conv." in key:  # ConvTranspose2d: PyTorch's fan_in uses dim 1
        return int(shape[1] * shape[2] * shape[3])
    if "gru.sequence_model" in key:
        return int(shape[0] // 3)  # nn.GRU: U(-1/sqrt(hidden), 1/sqrt(hidden))
    return int(np.prod(shape[1:]))


def make_crn_weights(seed: int = 0, gain: float = 1.0, **cfg):
    """Deterministic float32 parameters, PyTorch-default-like scales (U(+-1/sqrt(fan_in))).

    Norm weights are drawn around 1 and biases around 0 so that every affine term is exercised (default init of
    ones/zeros would hide a wrong per-channel index).  Returns an OrderedDict[str, np.ndarray].
    """
    shapes = crn_param_shapes(**cfg)
    out = OrderedDict()
    for key, shape in shapes.items():
        n = int(np.prod(shape))
        u = uniform01(seed, n, stream=zlib.crc32(key.encode())) * 2.0 - 1.0
        if "norm.weight" in key or "residualnorm.weight" in key:
            v = 1.0 + 0.25 * u
        elif "norm.bias" in key or "residualnorm.bias" in key:
            v = 0.1 * u
        else:
            fi = _fan_in(key, shape)
            if fi == 0:  # conv / linear / gru bias: bound by the matching weight's fan-in; 0.1 is representative
                v = 0.1 * u
            else:
                v = gain * u / np.sqrt(fi)
        out[key] = v.astype(np.float32).reshape(shape)
    return out


def with_alias_keys(weights):
    """Add the reference's ``net.0`` alias keys so the dict loads with strict=True into the reference module."""
    full = OrderedDict()
    for k, v in weights.items():
        full[k] = v
        if ".conv.weight" in k or ".conv.bias" in k:
            full[k.replace(".conv.", ".net.0.")] = v
    return full


# ----------------------------------------------------------------------------------------------------------------
# mixtures
# ----------------------------------------------------------------------------------------------------------------
def make_mixture(num_streams: int, length: int, num_mics: int = 3, first_stream: int = 0, sr: int = 16000):
    """Synthetic noisy multi-mic streams [B, M, L] float32 and the clean mic-0 source [B, L]."""
    mix = np.zeros((num_streams, num_mics, length), dtype=np.float32)
    src = np.zeros((num_streams, length), dtype=np.float32)
    t = np.arange(length + 8, dtype=np.float64) / sr
    for b in range(num_streams):
        s = first_stream + b
        seed = 2021 + s
        u = uniform01(seed, 16, stream=1)
        f0 = 100.0 + 200.0 * u[0]
        fenv = 3.0 + 2.0 * u[1]
        snr_db = -5.0 + 30.0 * u[2]
        amps = 0.2 + 0.8 * u[3:8]
        clean = np.zeros_like(t)
        for h in range(5):
            clean += amps[h] / (h + 1) * np.sin(2 * np.pi * (h + 1) * f0 * t + 2 * np.pi * u[8 + h])
        clean *= 0.5 * (1.0 + np.sin(2 * np.pi * fenv * t + 2 * np.pi * u[13]))
        clean *= 0.3 / (np.sqrt(np.mean(clean ** 2)) + 1e-12)
        p_clean = np.mean(clean ** 2)
        for m in range(num_mics):
            delayed = clean[8 - m * 2: 8 - m * 2 + length] if m > 0 else clean[8: 8 + length]
            noise = normal(seed, length, stream=100 + m)
            noise *= np.sqrt(p_clean / (10 ** (snr_db / 10.0)) / (np.mean(noise ** 2) + 1e-12))
            x = delayed + noise
            mix[b, m] = x.astype(np.float32)
        src[b] = clean[8: 8 + length].astype(np.float32)
        peak = np.abs(mix[b]).max()
        if peak > 0.95:
            mix[b] *= np.float32(0.95 / peak)
            src[b] *= np.float32(0.95 / peak)
    return mix, src


# ----------------------------------------------------------------------------------------------------------------
# FullSubNet weights (reference fullsubnet.py:729-747: two SequenceModels = nn.LSTM + nn.Linear)
# ----------------------------------------------------------------------------------------------------------------
def fsn_param_shapes(num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15, fb_num_neighbors=0,
                     num_layers=2):
    shapes = OrderedDict()

    def seq(prefix, isz, hidden, osz):
        for l in range(num_layers):
            shapes[f"{prefix}.sequence_model.weight_ih_l{l}"] = (4 * hidden, isz if l == 0 else hidden)
            shapes[f"{prefix}.sequence_model.weight_hh_l{l}"] = (4 * hidden, hidden)
            shapes[f"{prefix}.sequence_model.bias_ih_l{l}"] = (4 * hidden,)
            shapes[f"{prefix}.sequence_model.bias_hh_l{l}"] = (4 * hidden,)
        shapes[f"{prefix}.fc_output_layer.weight"] = (osz, hidden)
        shapes[f"{prefix}.fc_output_layer.bias"] = (osz,)

    seq("fb_model", num_freqs * num_mics, fb_hidden, num_freqs)
    seq("sb_model", (2 * sb_num_neighbors + 1) + (2 * fb_num_neighbors + 1), sb_hidden, 2)
    return shapes


def make_fsn_weights(seed: int = 0, **cfg):
    """Deterministic float32 FullSubNet parameters at PyTorch-default scales (U(+-1/sqrt(hidden)) / U(+-1/sqrt(fan_in)))."""
    out = OrderedDict()
    for key, shape in fsn_param_shapes(**cfg).items():
        n = int(np.prod(shape))
        u = uniform01(seed, n, stream=zlib.crc32(key.encode())) * 2.0 - 1.0
        if "sequence_model" in key:
            fi = shape[0] // 4
        else:
            fi = shape[-1] if key.endswith("weight") else cfg_hidden_of(key, cfg)
        out[key] = (u / np.sqrt(fi)).astype(np.float32).reshape(shape)
    return out


def cfg_hidden_of(key, cfg):
    return cfg.get("fb_hidden", 512) if key.startswith("fb_model") else cfg.get("sb_hidden", 384)
