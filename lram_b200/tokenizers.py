"""Action / return discretisers with the reference's interface and results
(`src/tokenizers_custom/__init__.py:1-14`, `minmax_tokenizer.py:5-79`, `mu_law_tokenizer.py:10-63`).

Same class names, constructor keywords (`vocab_size`, `shift`, `min_val`, `max_val`, `one_hot`) and
`tokenize` / `inv_tokenize` methods, so a policy written against the reference can import these instead.
On the rollout hot path the inverse of `MinMaxTokenizer` is additionally fused into the CUDA argmax
kernel (lram_b200/csrc/xl_elementwise.cu: argmax_tokens_kernel); `tests/test_oracle_pins.py` pins both to
known answers produced by the reference's own code (tests/golden/tokenizer_kat.npz).
"""
from __future__ import annotations

import math

import numpy as np
import torch


class BaseTokenizer:
    def __init__(self, vocab_size: int = 256, shift: int = 0):
        self._vocab_size = int(vocab_size)
        self._shift = int(shift)

    @property
    def vocab_size(self) -> int:
        return self._vocab_size

    @property
    def shift(self) -> int:
        return self._shift

    def tokenize(self, x):
        raise NotImplementedError

    def inv_tokenize(self, x):
        raise NotImplementedError


class MinMaxTokenizer(BaseTokenizer):
    """Uniform bins over [min_val, max_val): tok = clamp(trunc((x - min) / bin_width), 0, V-1) + shift."""

    def __init__(self, min_val=-1, max_val=1, one_hot=False, **kwargs):
        super().__init__(**kwargs)
        self.min_val, self.max_val, self.one_hot = min_val, max_val, one_hot
        self.bin_width = (max_val - min_val) / self.vocab_size

    def tokenize(self, x: torch.Tensor) -> torch.Tensor:
        # true division by bin_width then truncation toward zero (`.long()`), as the reference does
        ids = torch.div(x - self.min_val, self.bin_width).to(torch.long)
        ids = ids.clamp_(0, self.vocab_size - 1)
        if self.shift != 0:
            return ids + self.shift
        if self.one_hot:
            return torch.nn.functional.one_hot(ids, num_classes=self.vocab_size).float().flatten(-2)
        return ids

    def inv_tokenize(self, x: torch.Tensor) -> torch.Tensor:
        if self.one_hot:
            x = x.argmax(dim=-1)
        if self.shift != 0:
            x = (x - self.shift).clamp_min(0)       # ids below the shift decode to min_val
        return x.to(torch.float32) * self.bin_width + self.min_val


class MinMaxTokenizer2(BaseTokenizer):
    """RT-1 style: clamp, normalise to [0,1], scale by (V-1), truncate."""

    def __init__(self, min_val=-1, max_val=1, **kwargs):
        super().__init__(**kwargs)
        self.min_val, self.max_val = min_val, max_val

    def tokenize(self, x: torch.Tensor) -> torch.Tensor:
        span = self.max_val - self.min_val
        unit = (torch.clamp(x, self.min_val, self.max_val) - self.min_val) / span
        ids = (unit * (self.vocab_size - 1)).to(torch.long)
        return ids + self.shift if self.shift != 0 else ids

    def inv_tokenize(self, x: torch.Tensor) -> torch.Tensor:
        if self.shift != 0:
            x = (x - self.shift).clamp_min(0)
        unit = x.to(torch.float32) / (self.vocab_size - 1)
        return unit * (self.max_val - self.min_val) + self.min_val


class MuLawTokenizer(BaseTokenizer):
    """mu-law companding with mu = V-1 on signals in [-1, 1]."""

    def tokenize(self, x):
        mu = self.vocab_size - 1
        if isinstance(x, np.ndarray):
            comp = np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)
            ids = ((comp + 1) / 2 * mu + 0.5).astype(int)
        elif isinstance(x, torch.Tensor):
            xf = x.float()
            mu_t = torch.tensor([float(mu)], device=xf.device)
            comp = torch.sign(xf) * torch.log1p(mu_t * xf.abs()) / torch.log1p(mu_t)
            ids = ((comp + 1) / 2 * mu_t + 0.5).to(torch.long)
        else:
            raise NotImplementedError()
        return ids + self.shift if self.shift != 0 else ids

    def inv_tokenize(self, x_mu):
        mu = self.vocab_size - 1.0
        if self.shift != 0:
            x_mu = x_mu - self.shift
        if isinstance(x_mu, np.ndarray):
            y = (x_mu / mu) * 2 - 1.0
            return np.sign(y) * (np.exp(np.abs(y) * math.log1p(mu)) - 1.0) / mu
        if isinstance(x_mu, torch.Tensor):
            xf = x_mu.float()
            mu_t = torch.tensor([mu], device=xf.device)
            y = (xf / mu_t) * 2 - 1.0
            return torch.sign(y) * (torch.exp(y.abs() * torch.log1p(mu_t)) - 1.0) / mu_t
        raise NotImplementedError()


def make_tokenizer(kind, tokenizer_kwargs=None):
    kwargs = {} if tokenizer_kwargs is None else tokenizer_kwargs
    table = {"mulaw": MuLawTokenizer, "minmax": MinMaxTokenizer, "minmax2": MinMaxTokenizer2}
    if kind not in table:
        raise ValueError(f"Unknown tokenizer type {kind}")
    return table[kind](**kwargs)
