"""CPU oracle for the FullSubNet chunked streaming path (reference fullsubnet.py).

TEST INFRASTRUCTURE ONLY (see oracle/crn_oracle.py): a plain PyTorch-fp32 restatement used by tests/ and smoke();
the product package never imports it.  Pinned against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/fsn_*.npz).  On CPU the reference's ``with autocast()`` (fullsubnet.py:943) is a
CUDA-only context and leaves the arithmetic in fp32, which is what this file restates.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .crn_oracle import decompress_cirm, istft_chunk, over_add, segmentation, stft_chunk

EPS = 1e-8  # fullsubnet.py:22


def unfold(x: torch.Tensor, n: int) -> torch.Tensor:
    """BaseModel.unfold (fullsubnet.py:299-331): [B, C, F, T] -> [B, F, C, 2n+1, T] with reflect padding along F."""
    B, C, Fq, T = x.shape
    if n < 1:
        return x.permute(0, 2, 1, 3).reshape(B, Fq, C, 1, T)
    idx = torch.arange(Fq)[:, None] + torch.arange(2 * n + 1)[None, :] - n  # f + k - n
    idx = idx.abs()
    idx = torch.where(idx > Fq - 1, 2 * (Fq - 1) - idx, idx)  # reflect without repeating the edge
    out = x[:, :, idx, :]  # [B, C, F, 2n+1, T]
    return out.permute(0, 2, 1, 3, 4).contiguous()


class CumNorm:
    """CumLayerNorm (fullsubnet.py:177-205): running mean with alpha = step/(step+1), step capped at 80."""

    def __init__(self):
        self.mean, self.step = None, 0

    def __call__(self, x):
        dims = tuple(range(1, x.dim()))
        mean = x.mean(dim=dims, keepdim=True)
        if self.mean is None:
            self.mean = mean
        else:
            alpha = self.step / (self.step + 1)
            self.mean = alpha * self.mean + (1.0 - alpha) * mean
        self.step = min(self.step + 1, 80)
        x /= self.mean + EPS  # in place: the caller's tensor is normalised too (fullsubnet.py:200)
        return x


def lstm(x, state, weights, prefix, num_layers=2):
    """nn.LSTM (batch_first, gate order i,f,g,o) restated step by step.  x [N, T, I]; state (h, c) each [L, N, H]."""
    h0, c0 = state
    hs, cs = [], []
    inp = x
    for l in range(num_layers):
        wi, wh = weights[f"{prefix}.weight_ih_l{l}"], weights[f"{prefix}.weight_hh_l{l}"]
        bi, bh = weights[f"{prefix}.bias_ih_l{l}"], weights[f"{prefix}.bias_hh_l{l}"]
        h, c = h0[l], c0[l]
        outs = []
        for t in range(inp.shape[1]):
            g = inp[:, t] @ wi.T + bi + h @ wh.T + bh
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        hs.append(h)
        cs.append(c)
    return inp, (torch.stack(hs), torch.stack(cs))


class FSNOracle:
    def __init__(self, weights, num_freqs=201, num_mics=3, fb_hidden=512, sb_hidden=384, sb_num_neighbors=15,
                 fb_num_neighbors=0, num_layers=2, segment_length=3200, n_fft=400, hop=160):
        self.w = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in weights.items()}
        self.F, self.M, self.Hf, self.Hs = num_freqs, num_mics, fb_hidden, sb_hidden
        self.nb_sb, self.nb_fb, self.L = sb_num_neighbors, fb_num_neighbors, num_layers
        self.K, self.n_fft, self.hop = segment_length, n_fft, hop
        self.fh = self.sh = None
        self.norm_fb, self.norm_sb = CumNorm(), CumNorm()

    def reset_state(self, B):  # fullsubnet.py:826-832
        self.fh = (torch.zeros(self.L, B, self.Hf), torch.zeros(self.L, B, self.Hf))
        self.sh = (torch.zeros(self.L, B * self.F, self.Hs), torch.zeros(self.L, B * self.F, self.Hs))
        self.norm_fb, self.norm_sb = CumNorm(), CumNorm()

    def forward(self, x):
        """x [B, 2M, F, T] (real planes then imaginary planes) -> [B, 2, F, T]  (fullsubnet.py:769-824)."""
        M = self.M
        noisy = torch.sqrt(x[:, :M] ** 2 + x[:, M:] ** 2 + EPS)
        B, C, Fq, T = noisy.shape
        fb_in = self.norm_fb(noisy).reshape(B, C * Fq, T)  # `noisy` itself is now normalised (in-place)
        o, self.fh = lstm(fb_in.permute(0, 2, 1), self.fh, self.w, "fb_model.sequence_model", self.L)
        o = torch.relu(o @ self.w["fb_model.fc_output_layer.weight"].T + self.w["fb_model.fc_output_layer.bias"])
        fb_out = o.permute(0, 2, 1).unsqueeze(1)  # [B, 1, F, T]
        fb_unf = unfold(fb_out, self.nb_fb).reshape(B, Fq, 2 * self.nb_fb + 1, T)
        noisy_unf = unfold(noisy[:, 0].unsqueeze(1), self.nb_sb).reshape(B, Fq, 2 * self.nb_sb + 1, T)
        sb_in = self.norm_sb(torch.cat([noisy_unf, fb_unf], dim=2))
        sb_in = sb_in.reshape(B * Fq, -1, T)
        o, self.sh = lstm(sb_in.permute(0, 2, 1), self.sh, self.w, "sb_model.sequence_model", self.L)
        o = o @ self.w["sb_model.fc_output_layer.weight"].T + self.w["sb_model.fc_output_layer.bias"]  # [B*F, T, 2]
        return o.permute(0, 2, 1).reshape(B, Fq, 2, T).permute(0, 2, 1, 3).contiguous()

    def stft_trans(self, chunks):
        """[R, M, K] -> [R, 2M, F, T]  (fullsubnet.py:835-844)."""
        R, M, K = chunks.shape
        s = stft_chunk(chunks.reshape(R * M, K), self.n_fft, self.hop)  # [R*M, T, F, 2]
        s = s.reshape(R, M, -1, self.F, 2).transpose(2, 3)              # [R, M, F, T, 2]
        return torch.cat([s[..., 0], s[..., 1]], dim=1)

    def realtime_process(self, mixture, flag=False):
        """The train=False chunk loop of fullsubnet.py:903-961 -> (pred [B, L], pred_crm [N, B, 2, F, T])."""
        B, C, _ = mixture.shape
        P = self.K // 2
        if not flag:
            mixture = torch.cat([torch.zeros(B, C, P), mixture], dim=-1)
        seg, gap = segmentation(mixture, self.K)
        N = seg.shape[0] // B
        x = self.stft_trans(seg).reshape(B, N, 2 * C, self.F, -1).transpose(0, 1)  # [N, B, 2M, F, T]
        if not flag:
            self.reset_state(B)
        crm = torch.stack([self.forward(x[n].clone()) for n in range(N)])  # [N, B, 2, F, T]
        m = decompress_cirm(crm)
        xr, xi = x[:, :, 0], x[:, :, C]
        er = m[:, :, 0] * xr - m[:, :, 1] * xi
        ei = m[:, :, 1] * xr + m[:, :, 0] * xi
        e = torch.stack([er, ei], dim=-1)  # [N, B, F, T, 2]
        y = istft_chunk(e.reshape(N * B, self.F, -1, 2).permute(0, 2, 1, 3), self.n_fft, self.hop)
        y = over_add(y.reshape(N, B, -1).permute(1, 0, 2), gap)
        return (y if flag else y[..., P:]), crm
