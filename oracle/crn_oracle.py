"""CPU oracle for the CRN_ELU streaming-enhancement hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch-fp32 (CPU) restatement of the reference algorithm; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path (``speech_enhancement_mi_b200``) never does.

Pinning: the reference ships no tests or golden vectors, so parity is pinned by executing the
UNMODIFIED reference files in the build container (``oracle/make_golden.py`` under ``oracle/shim``) and committing the
resulting vectors to ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPS = 1e-8  # CRN_ELU.py:11


# ---------------------------------------------------------------------------------------------------------------
# a1: padding / segmentation  (utility.py:312-370)  -- integer index math, bit-exact
# ---------------------------------------------------------------------------------------------------------------
def chunk_grid(length: int, K: int):
    """gap and chunk count for a signal of ``length`` samples entering segmentation (utility.py:325-327,357-368)."""
    P = K // 2
    gap = K - (P + length % K) % K
    n_chunks = 2 * (length + gap + P) // K
    return gap, n_chunks


def segmentation(x: torch.Tensor, K: int):
    """[B, C, L] -> ([B*N, C, K], gap): chunk n of stream b is row b*N+n and covers padded[n*P : n*P+K]."""
    B, C, L = x.shape
    P = K // 2
    gap, N = chunk_grid(L, K)
    padded = torch.zeros(B, C, P + L + gap + P, dtype=x.dtype)
    padded[..., P:P + L] = x
    out = torch.empty(B, N, C, K, dtype=x.dtype)
    for n in range(N):
        out[:, n] = padded[..., n * P:n * P + K]
    return out.reshape(B * N, C, K), gap


def over_add(chunks: torch.Tensor, gap: int):
    """[C, N, K] -> [C, L]: average of the even-chunk and odd-chunk tilings (utility.py:393-403)."""
    C, N, K = chunks.shape
    P = K // 2
    total = N * P + P  # padded grid length
    acc = torch.zeros(C, total, dtype=chunks.dtype)
    for n in range(N):
        acc[:, n * P:n * P + K] += chunks[:, n]
    y = acc[:, P:total - P] / 2
    return y[:, :y.shape[1] - gap] if gap > 0 else y


# ---------------------------------------------------------------------------------------------------------------
# a2 / a10: STFT / iSTFT (CRN_ELU.py:329-333,417-432 -> speechbrain STFT/ISTFT -> torch.stft/istft)
# Restated as explicit framing + rFFT so that framing indices are visible.
# ---------------------------------------------------------------------------------------------------------------
def hamming_periodic(n: int) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64)
    return (0.54 - 0.46 * torch.cos(2 * math.pi * k / n)).to(torch.float32)


def stft_chunk(x: torch.Tensor, n_fft=400, hop=160) -> torch.Tensor:
    """[R, K] -> [R, T, F, 2]; zero-pad n_fft/2 both sides, frame t = padded[hop*t : hop*t+n_fft] * hamming."""
    R, K = x.shape
    half = n_fft // 2
    padded = F.pad(x, (half, half))
    T = 1 + K // hop
    idx = (torch.arange(T) * hop)[:, None] + torch.arange(n_fft)[None, :]
    frames = padded[:, idx] * hamming_periodic(n_fft)
    spec = torch.fft.rfft(frames, n=n_fft, dim=-1)
    return torch.view_as_real(spec)


def istft_chunk(spec: torch.Tensor, n_fft=400, hop=160) -> torch.Tensor:
    """[R, T, F, 2] -> [R, hop*(T-1)]: irFFT, window, overlap-add, divide by the window-square envelope, trim."""
    R, T, Fq, _ = spec.shape
    w = hamming_periodic(n_fft)
    frames = torch.fft.irfft(torch.view_as_complex(spec.contiguous()), n=n_fft, dim=-1) * w
    full = n_fft + hop * (T - 1)
    y = torch.zeros(R, full, dtype=spec.dtype)
    env = torch.zeros(full, dtype=spec.dtype)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, t]
        env[t * hop:t * hop + n_fft] += w * w
    half = n_fft // 2
    return y[:, half:full - half] / env[half:full - half]


# ---------------------------------------------------------------------------------------------------------------
# a6: GlobalLayerNorm (CRN_ELU.py:37-56; student denominator distillation_crn.py:51)
# ---------------------------------------------------------------------------------------------------------------
def gln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, student=False) -> torch.Tensor:
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    den = (torch.sqrt(var) + EPS) if student else (torch.sqrt(var + EPS) + EPS)
    return (x - mean) / den * w + b


# ---------------------------------------------------------------------------------------------------------------
# a9: decompress_cIRM (utility.py:439-442)
# ---------------------------------------------------------------------------------------------------------------
def decompress_cirm(m: torch.Tensor, K=10.0, limit=9.9) -> torch.Tensor:
    m = limit * (m >= limit) - limit * (m <= -limit) + m * (m.abs() < limit)
    return -K * torch.log((K - m) / (K + m))


class CRNOracle:
    """Functional, explicitly-stateful restatement of ``TemporalCRN`` (CRN_ELU.py:314-509).

    ``weights``: dict key -> torch.float32 tensor with the reference state_dict keys (alias keys optional).
    ``student=True`` selects the numerics of distillation_crn.TemporalCRN (distillation_crn.py:51,340).
    """

    def __init__(self, weights, num_channels, num_freqs=201, hidden=512, segment_length=3200, num_layers=2,
                 num_inputs=3, kernel_size=3, sample_rate=16000, win_length=25, hop_length=10, n_fft=400,
                 student=False, **_unused):
        self.w = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in weights.items()}
        self.num_channels = list(num_channels)
        self.L = len(self.num_channels)
        self.num_freqs, self.hidden, self.K = num_freqs, hidden, segment_length
        self.num_layers, self.num_inputs, self.kernel_size = num_layers, num_inputs, kernel_size
        self.n_fft = n_fft
        self.win = int(round(sample_rate / 1000.0 * win_length))
        self.hop = int(round(sample_rate / 1000.0 * hop_length))
        assert self.win == n_fft, "reference config uses win_length == n_fft"
        self.student = student
        self.reset()

    # -- state (CRN_ELU.py:158,228,408-415) -------------------------------------------------------------------
    def reset(self):
        self.buf = {}
        self.h = None

    # distillation_crn.py:343-377: when a list, forward() appends the five pre-activation `feature` tensors to it
    taps = None

    # -- a5: TemporalConv2d (CRN_ELU.py:230-247) --------------------------------------------------------------
    def _tconv(self, name, x, stride, dilation, pad_f, pad_t):
        w = self.w
        B, C, Fq, T = x.shape
        state = self.buf.get(name)
        if state is None:
            state = torch.zeros(B, C, Fq, pad_t)
        inp = torch.cat([state, x], dim=-1)
        o = F.conv2d(inp, w[f"{name}.conv.weight"], w[f"{name}.conv.bias"], stride=stride, padding=(pad_f, 0),
                     dilation=dilation)
        self._feature = o  # distillation_crn.py:198 `feature = self.net(inp)`
        o = F.elu(o)
        o = F.conv2d(o, w[f"{name}.conv_trans.weight"], w[f"{name}.conv_trans.bias"]) * torch.sigmoid(
            F.conv2d(o, w[f"{name}.conv_gated.weight"], w[f"{name}.conv_gated.bias"]))
        o = gln(o, w[f"{name}.norm.weight"], w[f"{name}.norm.bias"], self.student)
        assert T > pad_t
        self.buf[name] = x[..., -pad_t:].detach().clone()  # CRN_ELU.py:243: detached
        return o

    # -- a8: TemporalConvTranspose2d (CRN_ELU.py:290-307) -----------------------------------------------------
    def _tdeconv(self, name, x, dilation_t, res=None):
        w = self.w
        T = x.shape[-1]
        o = F.conv_transpose2d(x, w[f"{name}.conv.weight"], w[f"{name}.conv.bias"], stride=(2, 1), padding=(2, 0),
                               dilation=(1, dilation_t))[..., -T:]
        self._feature = o  # distillation_crn.py:253 `feature = out`
        o = F.elu(o)
        o = gln(o, w[f"{name}.norm.weight"], w[f"{name}.norm.bias"], self.student)
        if res is not None:
            Fr = res.shape[2]
            if Fr > o.shape[2]:
                o = F.pad(o, (0, 0, 0, Fr - o.shape[2]))
            elif Fr < o.shape[2]:
                o = o[:, :, :Fr]
            m = torch.sigmoid(gln(F.conv2d(res, w[f"{name}.residualmask.weight"], w[f"{name}.residualmask.bias"]),
                                  w[f"{name}.residualnorm.weight"], w[f"{name}.residualnorm.bias"], self.student))
            r = F.elu(F.conv2d(res, w[f"{name}.residual.weight"], w[f"{name}.residual.bias"]))
            o = m * r + (1.0 - m) * o
        return o

    # -- a7: SequenceModel = GRU + Linear + ELU + GLN(last) (CRN_ELU.py:160-186) -------------------------------
    def _gru(self, x):
        w = self.w
        B, Fe, T = x.shape
        H = self.hidden
        seq = x.permute(0, 2, 1)  # [B, T, Fe]
        if self.h is None:
            self.h = torch.zeros(self.num_layers, B, H)
        h_out = []
        for l in range(self.num_layers):
            wih, whh = w[f"gru.sequence_model.weight_ih_l{l}"], w[f"gru.sequence_model.weight_hh_l{l}"]
            bih, bhh = w[f"gru.sequence_model.bias_ih_l{l}"], w[f"gru.sequence_model.bias_hh_l{l}"]
            gi = seq @ wih.t() + bih  # [B, T, 3H], gate order r, z, n
            h = self.h[l]
            outs = []
            for t in range(T):
                gh = h @ whh.t() + bhh
                r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
                h = (1.0 - z) * n + z * h
                outs.append(h)
            seq = torch.stack(outs, dim=1)
            h_out.append(h)
        self.h = torch.stack(h_out, dim=0).detach()  # CRN_ELU.py:185: detached
        o = seq @ w["gru.fc_output_layer.weight"].t() + w["gru.fc_output_layer.bias"]
        self._feature = o  # distillation_crn.py:140 `feature = o`, [B, T, C*F]
        o = F.elu(o)
        o = gln(o.unsqueeze(1), w["gru.norm.weight"], w["gru.norm.bias"], self.student).squeeze(1)
        return o.permute(0, 2, 1)

    # -- a4: input features (CRN_ELU.py:369-373; student distillation_crn.py:340) ------------------------------
    def features(self, x):
        re, im = x[..., 0], x[..., 1]
        if self.student:
            angle = torch.arctan(im / (re + EPS) + EPS)
        else:
            angle = torch.atan2(im, re)
        ipd = angle[:, :1] - angle[:, 1:]
        mag = torch.sqrt(re ** 2 + im ** 2 + 1e-10)
        return torch.cat([mag, ipd], dim=1)

    # -- forward (CRN_ELU.py:367-406) ------------------------------------------------------------------------
    def forward(self, x, return_mask=False, trace=None):
        """x [B, M, F, T, 2] (STFT of one chunk) -> enhanced spectrum [B, F, T, 2].

        ``trace``: optional dict that receives the activation entering every block ([B, C, F, T]), keyed like the
        CUDA path's internal buffers (pre_in<i>, enc_in<i>, xg, dec_in<j>) -- used to localise parity failures."""
        noisy = x[:, 0]
        y = self.features(x)
        for i, d in enumerate((1, 2, 4)):
            if trace is not None:
                trace[f"pre_in{i}"] = y
            y = self._tconv(f"preconvlist.{i}", y, (1, 1), (d, 1), 2 * d, 4) + y
        residuals = [y]
        for i in range(self.L):
            d = 2 ** i
            if trace is not None:
                trace[f"enc_in{i}"] = y
            y = self._tconv(f"convlist.{i}", y, (2, 1), (1, d), 2, (self.kernel_size - 1) * d)
            residuals.append(y)
        B, C, Fq, T = y.shape
        taps = self.taps
        if taps is not None:
            taps.append(self._feature)  # of the LAST encoder block only (distillation_crn.py:355)
        if trace is not None:
            trace["xg"] = y
        y = self._gru(y.reshape(B, C * Fq, T)).reshape(B, C, Fq, T)
        if taps is not None:
            taps.append(self._feature.reshape(B, C, Fq, T))  # a plain re-shape of [B, T, C*F] (distillation_crn.py:364)
        idx = -2
        for j in range(self.L - 1):
            if trace is not None:
                trace[f"dec_in{j}"] = y
            y = self._tdeconv(f"deconvlist.{j}", y, 2 ** j, residuals[idx])
            if taps is not None:
                taps.append(self._feature)
            idx -= 1
        if trace is not None:
            trace[f"dec_in{self.L - 1}"] = y
        y = self._tdeconv(f"deconvlist.{self.L - 1}", y, 2 ** (self.L - 1)).permute(0, 2, 3, 1)
        m = decompress_cirm(y)
        er = m[..., 0] * noisy[..., 0] - m[..., 1] * noisy[..., 1]
        ei = m[..., 1] * noisy[..., 0] + m[..., 0] * noisy[..., 1]
        out = torch.stack([er, ei], dim=-1)
        return (out, y) if return_mask else out

    # -- per-chunk transforms (CRN_ELU.py:417-432) -------------------------------------------------------------
    def stft_trans(self, chunks):
        """[R, M, K] -> [R, M, F, T, 2]."""
        R, M, K = chunks.shape
        s = stft_chunk(chunks.reshape(R * M, K), self.n_fft, self.hop)
        return s.reshape(R, M, -1, self.num_freqs, 2).transpose(2, 3)

    def istft_trans(self, spec):
        """[R, F, T, 2] -> [R, K]."""
        return istft_chunk(spec.permute(0, 2, 1, 3), self.n_fft, self.hop)

    # -- a11: realtime_process (CRN_ELU.py:472-509) -------------------------------------------------------------
    def realtime_process(self, mixture, flag=False, return_features=False):
        """return_features: also return the five feature taps, each [N*B, C, F, T] with the chunks concatenated along
        the batch axis (distillation_crn.py:454-470)."""
        B, C, _ = mixture.shape
        P = self.K // 2
        if not flag:
            mixture = torch.cat([torch.zeros(B, C, P), mixture], dim=-1)
            self.reset()
        seg, gap = segmentation(mixture, self.K)
        N = seg.shape[0] // B
        spec = self.stft_trans(seg).reshape(B, N, C, self.num_freqs, -1, 2)
        outs, feats = [], []
        for n in range(N):
            self.taps = [] if return_features else None
            e = self.forward(spec[:, n])
            if return_features:
                feats.append(self.taps)
            self.taps = None
            outs.append(self.istft_trans(e))
        y = over_add(torch.stack(outs, dim=1), gap)
        y = y if flag else y[..., P:]
        if return_features:
            return y, [torch.cat([f[i] for f in feats], dim=0) for i in range(len(feats[0]))]
        return y

    # -- the true-streaming step the CUDA path exposes: chunk in -> hop samples out (SURVEY.md section 3.1 probe) ---
    def stream_step(self, chunk, carry):
        """chunk [B, M, K]; carry [B, K/2] or None.  Returns (out [B, K/2], new carry [B, K/2])."""
        y = self.istft_trans(self.forward(self.stft_trans(chunk)))
        P = self.K // 2
        if carry is None:
            carry = torch.zeros(chunk.shape[0], P)
        return (y[:, :P] + carry) / 2, y[:, P:].clone()


# ---------------------------------------------------------------------------------------------------------------
# a13 (part): SI-SNR term (utility.py:207-223) and the SI-SDR of metrics.py:61-85 used for the parity "SI-SDR delta"
# ---------------------------------------------------------------------------------------------------------------
def cal_si_snr(separated, source, length=None, eps=1e-8):
    total = 0.0
    B = separated.shape[0]
    for i in range(B):
        n = separated.shape[1] if length is None else int(length[i])
        a = separated[i, :n] - separated[i, :n].mean()
        s = source[i, :n] - source[i, :n].mean()
        proj = (a * s).sum() * s / (s.norm() ** 2 + eps)
        total = total + 20 * torch.log10(eps + proj.norm() / ((a - proj).norm() + eps))
    return total / B


def si_sdr_db(estimate, reference):
    """Scale-invariant SDR in dB of ``estimate`` w.r.t. ``reference`` (metrics.py:61-85), float64."""
    e = torch.as_tensor(estimate, dtype=torch.float64).flatten()
    r = torch.as_tensor(reference, dtype=torch.float64).flatten()
    e = e - e.mean()
    r = r - r.mean()
    alpha = (e * r).sum() / ((r * r).sum() + 1e-30)
    target = alpha * r
    noise = e - target
    return float(10 * torch.log10((target * target).sum() / ((noise * noise).sum() + 1e-30)))


# ---------------------------------------------------------------------------------------------------------------
# a13: STOI-like loss (utility.py:821-916 with thirdoct :480-518 and removeSilentFrames :521-571), restated with the
# same torch / torchaudio CPU operators the reference calls
# ---------------------------------------------------------------------------------------------------------------
def thirdoct(fs=10000, nfft=512, num_bands=15, min_freq=150):
    import numpy as np
    f = torch.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = torch.arange(num_bands, dtype=torch.float64)
    lo = min_freq * torch.pow(2.0, (2 * k - 1) / 6)
    hi = min_freq * torch.pow(2.0, (2 * k + 1) / 6)
    obm = torch.zeros(num_bands, len(f))
    for i in range(num_bands):
        a = int(torch.argmin(torch.square(f - lo[i])))
        b = int(torch.argmin(torch.square(f - hi[i])))
        obm[i, a:b] = 1
    return obm


def remove_silent_frames(x, y, dyn_range=40, N=256):
    import numpy as np
    w = torch.from_numpy(np.hanning(256)).to(torch.float)
    nf = x.shape[0] // N + (x.shape[0] - 128) // N
    starts = [128 * m for m in range(nf)]
    X = torch.stack([x[s:s + N] for s in starts], dim=1)
    Y = torch.stack([y[s:s + N] for s in starts], dim=1)
    energy = 20 * torch.log10(torch.sqrt((w[:, None] ** 2 * X ** 2).sum(0)) / 16.0 + np.finfo("float").eps)
    msk = energy - energy.max() + dyn_range > 0
    xs, ys = w[:, None] * X[:, msk], w[:, None] * Y[:, msk]

    def ola(z):
        return torch.cat((z[:128, 0], (z[:128, 1:] + z[128:, :-1]).T.flatten(), z[128:, -1]))

    return ola(xs), ola(ys)


def stoi_loss(y_true, y_pred, lens):
    import numpy as np
    import torchaudio
    eps = np.finfo("float").eps
    obm, c, Nseg = thirdoct(), 5.62341325, 30
    resampler = torchaudio.transforms.Resample(16000, 10000)
    spec = torchaudio.transforms.Spectrogram(n_fft=512, win_length=256, hop_length=128, power=2)
    D = torch.zeros(y_true.shape[0])
    for i in range(y_true.shape[0]):
        t, p = resampler(y_true[i, :int(lens[i])]), resampler(y_pred[i, :int(lens[i])])
        try:
            t, p = remove_silent_frames(t, p)
        except Exception:
            pass
        if t.shape[-1] <= 512:
            D[i] = 0.99
            continue
        ot, op = torch.sqrt(obm @ spec(t) + 1e-14), torch.sqrt(obm @ spec(p) + 1e-14)
        M = ot.shape[-1] - (Nseg - 1)
        if M <= 0:
            X, Y, M = ot, op, 1
        else:
            X = torch.cat([ot[:, m:m + Nseg] for m in range(M)], dim=0)
            Y = torch.cat([op[:, m:m + Nseg] for m in range(M)], dim=0)
        alpha = X.norm(dim=-1, keepdim=True) / (Y.norm(dim=-1, keepdim=True) + eps)
        y = torch.min(Y * alpha, X + X * c)
        xn = X - X.mean(-1, keepdim=True)
        xn = xn / (xn.norm(dim=-1, keepdim=True) + eps)
        yn = y - y.mean(-1, keepdim=True)
        yn = yn / (yn.norm(dim=-1, keepdim=True) + eps)
        D[i] = (xn * yn).sum() / (15.0 * M)
    return -D.mean()


def compute_loss(source, pred, length):
    """CRN_ELU.py:513-535 (without the print): (loss, mae, sisnr)."""
    mae = stoi_loss(source, pred, length)
    sisnr = -cal_si_snr(pred, source, length)
    return 0.7 * mae + 0.3 * sisnr, mae, sisnr


# ---------------------------------------------------------------------------------------------------------------
# f4: distillation loss (distillation_crn.py:548-564).  connectors[i] = (conv weight [Ct, Cs, 1, 1], bn weight [Ct],
# bn bias [Ct]); BatchNorm2d in training mode = batch statistics over (N, F, T), biased variance, eps 1e-5.
# ---------------------------------------------------------------------------------------------------------------
def distillation_loss(ft, fs, connectors):
    loss = 0.0
    for t, s, (cw, bw, bb) in zip(ft, fs, connectors):
        neg = (t < 0.0).float()
        margin = (t * neg).sum(dim=(0, 2, 3), keepdim=True) / (neg.sum(dim=(0, 2, 3), keepdim=True) + EPS)
        t = torch.max(t, margin)
        s = F.conv2d(s, cw)
        mean = s.mean(dim=(0, 2, 3), keepdim=True)
        var = s.var(dim=(0, 2, 3), keepdim=True, unbiased=False)
        s = (s - mean) / torch.sqrt(var + 1e-5) * bw.view(1, -1, 1, 1) + bb.view(1, -1, 1, 1)
        mask = 1.0 - ((s <= t) & (t <= 0.0)).float()
        loss = loss + torch.mean((s - t) ** 2 * mask)
    return loss / len(ft)


# ---------------------------------------------------------------------------------------------------------------
# training micro-step (train.py:195-198): realtime_process -> compute_loss -> backward, by autograd through the
# restatement above.  The carried conv buffers / GRU state are detached exactly where the reference detaches them
# (CRN_ELU.py:185,243), so the gradient of every chunk stops at its own inputs.
# ---------------------------------------------------------------------------------------------------------------
def train_step_grads(oracle: "CRNOracle", mixture, source, lens, flag=False):
    """Returns (pred [B,L], d loss/d pred [B,L], (loss, mae, sisnr) floats, {name: grad}) for the weights of ``oracle``."""
    leaves = {}
    for k, v in oracle.w.items():
        if ".net.0." in k:
            continue
        leaves[k] = v.detach().clone().requires_grad_(True)
    saved = oracle.w
    oracle.w = dict(leaves)
    for k in saved:
        if ".net.0." in k:
            oracle.w[k] = leaves[k.replace(".net.0.", ".conv.")]
    try:
        pred = oracle.realtime_process(torch.as_tensor(mixture, dtype=torch.float32), flag)
        pred.retain_grad()
        loss, mae, sisnr = compute_loss(torch.as_tensor(source, dtype=torch.float32), pred, torch.as_tensor(lens))
        loss.backward()
    finally:
        oracle.w = saved
    grads = {k: v.grad.detach() for k, v in leaves.items() if v.grad is not None}
    return pred.detach(), pred.grad.detach(), (float(loss), float(mae), float(sisnr)), grads
