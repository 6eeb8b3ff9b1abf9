"""Restatement of speechbrain 0.5.x ``STFT`` / ``ISTFT`` (un-vendored, un-pinned: reference requirements.txt:14).

TEST INFRASTRUCTURE ONLY.  Used solely by ``oracle/make_golden.py`` so that the unmodified reference files under
``/root/reference`` import in the build container.  Only the 2-D ``[batch, time]`` input case is restated, which is the
only case the reference uses (CRN_ELU.py:421-423, fullsubnet.py:841,851).  The published behaviour of those classes:
milliseconds -> samples by ``round(sr/1000*ms)``, ``torch.hamming_window`` (periodic), ``torch.stft`` with
``center=True, pad_mode="constant"``, output transposed to ``[B, T, F, 2]``; ``torch.istft`` on the inverse layout.
"""
import torch


class STFT(torch.nn.Module):
    def __init__(self, sample_rate, win_length=25, hop_length=10, n_fft=400, window_fn=torch.hamming_window,
                 normalized_stft=False, center=True, pad_mode="constant", onesided=True):
        super().__init__()
        self.n_fft = n_fft
        self.win_length = int(round(sample_rate / 1000.0 * win_length))
        self.hop_length = int(round(sample_rate / 1000.0 * hop_length))
        self.normalized_stft, self.center, self.pad_mode, self.onesided = normalized_stft, center, pad_mode, onesided
        self.window = window_fn(self.win_length)

    def forward(self, x):
        s = torch.stft(x, self.n_fft, self.hop_length, self.win_length, self.window.to(x.device), self.center,
                       self.pad_mode, self.normalized_stft, self.onesided, return_complex=True)
        return torch.view_as_real(s).transpose(2, 1)


class ISTFT(torch.nn.Module):
    def __init__(self, sample_rate, n_fft=None, win_length=25, hop_length=10, window_fn=torch.hamming_window,
                 normalized_stft=False, center=True, onesided=True, epsilon=1e-12):
        super().__init__()
        self.n_fft = n_fft
        self.win_length = int(round(sample_rate / 1000.0 * win_length))
        self.hop_length = int(round(sample_rate / 1000.0 * hop_length))
        self.normalized_stft, self.center, self.onesided = normalized_stft, center, onesided
        self.window = window_fn(self.win_length)

    def forward(self, x, sig_length=None):
        x = x.permute(0, 2, 1, 3)
        x = torch.complex(x[..., 0].contiguous(), x[..., 1].contiguous())
        return torch.istft(x, n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length,
                           window=self.window.to(x.device), center=self.center, onesided=self.onesided,
                           length=sig_length)
