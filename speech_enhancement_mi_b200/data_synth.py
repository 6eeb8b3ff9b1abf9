"""Synthetic stand-in for the reference's ``data_c.LibriPartyDataset`` (private corpus, config.yaml:35-47): the same
item dict ``{'mix', 'source', 'noise', 'length', 'flag'}`` (data_c.py:60-83) and the same piece buffering -- a long
mixture is cut into pieces of random length in [16000, max_length) (data_c.py:156-175); the first piece of a mixture has
``flag=False``, the following pieces ``flag=True`` and continue the model state (CRN_ELU.py:474-481).

Signals come from ``synth.make_mixture`` (16 kHz, 3 microphones, SNR in [-5, 25] dB).  Host
logic only.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from . import synth


class SyntheticPartyDataset(Dataset):
    def __init__(self, size=64, utterance_seconds=6.0, max_length=60000, num_mic=3, seed=0, sample_rate=16000):
        self.size, self.num_mic, self.max_length = int(size), int(num_mic), int(max_length)
        self.utt_len = int(utterance_seconds * sample_rate)
        self.rng = np.random.RandomState(seed)
        self.buffer = []
        self.next_stream = 0

    def init_seed(self, seed):  # data_c.py:85-89
        self.rng = np.random.RandomState(seed)
        self.next_stream = 1000 * int(seed)

    def set_attribute(self, dataset="train", **_unused):  # data_c.py:25-52 (augmentation switches do not apply)
        self.buffer = []

    def __len__(self):
        return self.size

    def _fill(self):
        mix, src = synth.make_mixture(1, self.utt_len, self.num_mic, first_stream=self.next_stream)
        self.next_stream += 1
        mix, src = torch.from_numpy(mix[0]), torch.from_numpy(src[0])
        noise = mix - src[None, :]
        start = 0
        while start < self.utt_len:  # data_c.py:164-175 (with the evident intent `start = end`)
            piece = int(self.rng.randint(16000, self.max_length))
            end = min(self.utt_len, start + piece)
            if end - start < 16000:
                break
            self.buffer.append((mix[:, start:end], src[None, None, start:end], noise[:, start:end],
                                torch.tensor([end - start]), start > 0))
            start = end

    def __getitem__(self, index):
        while not self.buffer:
            self._fill()
        mix, source, noise, length, flag = self.buffer.pop(0)
        return {"mix": mix, "source": source, "noise": noise, "length": length, "flag": flag}
