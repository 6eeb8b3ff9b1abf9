"""Host-side description of the streaming workload: algorithmic FLOPs / bytes per stream-chunk of every stage of the
CRN path (2*MACs counted as the reference's module hooks would), and stream sharding."""
from __future__ import annotations

T_FRAMES = 21
N_FREQ = 201
CHUNK = 3200
HOP = CHUNK // 2
AUDIO_SEC_PER_STEP = HOP / 16000.0  # every step advances each stream by 0.1 s (utility.py:360-367)

TEACHER = dict(num_channels=[16, 32, 64, 128], num_freqs=201, hidden=512, num_layers=2, num_inputs=3, kernel_size=3)
STUDENT = dict(num_channels=[16, 32, 64, 64], num_freqs=201, hidden=128, num_layers=2, num_inputs=3, kernel_size=3)


def freq_pyramid(num_levels, F=N_FREQ):
    out = []
    for _ in range(num_levels):
        F = (F - 1) // 2 + 1
        out.append(F)
    return out


def algorithmic_flops(num_channels, hidden, num_inputs=3, kernel_size=3, num_layers=2, **_):
    """FLOPs (2*MACs) per stream per 3200-sample chunk, by stage (CRN_ELU.py:337-365 shapes)."""
    T, F = T_FRAMES, N_FREQ
    c0 = 2 * num_inputs - 1
    L = len(num_channels)
    fl = {}
    fl["preconv"] = 3 * (2 * T * F * c0 * (c0 * 25) + 2 * 2 * T * F * c0 * c0)
    pyr = freq_pyramid(L)
    enc = 0
    cin = c0
    for i, co in enumerate(num_channels):
        enc += 2 * T * pyr[i] * co * (cin * 5 * kernel_size) + 2 * 2 * T * pyr[i] * co * co
        cin = co
    fl["encoder"] = enc
    feat = pyr[-1] * num_channels[-1]
    gru = 0
    for l in range(num_layers):
        gru += 2 * T * 3 * hidden * ((feat if l == 0 else hidden) + hidden)
    gru += 2 * T * hidden * feat
    fl["gru"] = gru
    dec = 0
    fin = pyr[-1]
    for j in range(L):
        ci = num_channels[L - 1 - j]
        co = num_channels[L - 2 - j] if j < L - 1 else 2
        dec += 2 * T * fin * ci * co * 5 * kernel_size  # ConvTranspose2d, counted over its full output
        if j < L - 1:
            fs = pyr[L - 2 - j]
            dec += 2 * 2 * T * fs * co * co  # residualmask + residual 1x1
            fin = fs
    fl["decoder"] = dec
    fl["total"] = sum(fl.values())
    return fl


def algorithmic_bytes():
    """HBM bytes per stream-chunk of the bandwidth-bound stages."""
    T, F = T_FRAMES, N_FREQ
    stft = 3 * CHUNK * 4 + 5 * F * T * 4 + F * T * 2 * 4  # read chunk; write 5 feature planes + mic-0 spectrum
    mask = 2 * (F * T * 2 * 4) + HOP * 4 + 2 * HOP * 4    # read mask + spectrum; write hop; read+write carry
    return {"stft": stft, "mask_istft": mask}


def shard_streams(total_streams, world_size, rank):
    """Contiguous block of streams owned by `rank` (streams never interact: CRN_ELU.py:40-41 norms are per sample)."""
    base, rem = divmod(total_streams, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)
