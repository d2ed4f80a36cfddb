"""`xlstm` STAND-IN for running the reference's own model code — test infrastructure, NOT product code.

The reference imports the third-party `xlstm` package (`src/algos/models/decision_xlstm.py:8-12`), which is absent
from /root/reference and not installable here. This module offers the symbols those imports name, as `nn.Module`s
that carry xlstm v1.0.x's parameter names and whose arithmetic is the CPU oracle's
(`oracle/xlstm_oracle.py`): `xLSTMBlockStack.step` / `.forward`, `xLSTMBlockStackConfig` (+ nested configs, the
dict layout of `configs/agent_params/huggingface/xlstm_*.yaml:7-26`), `LayerNorm`, `MultiHeadLayerNorm`,
`LinearHeadwiseExpand`, `mLSTMCell`, `sLSTMCell_cuda`.

`tests/golden/ref_stubs.py` registers it in `sys.modules` as `xlstm` so that the reference's
`MultiDomainDiscreteDecisionXLSTMModel`, `DiscreteDecisionXLSTM.predict` and `custom_evaluate_policy` run unmodified;
what those runs pin is the LRAM-side logic (embedding, token interleave, cache trimming, head slicing, argmax,
inv_tokenize, rollout bookkeeping). The xlstm arithmetic itself stays pinned only by the `transformers` vectors
(see the oracle's header).
"""
from __future__ import annotations

import dataclasses
import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import xlstm_oracle as O


# ---- configs (dict layout read by dacite at decision_xlstm.py:130-132) ---------------------------------------------
@dataclass
class mLSTMLayerConfig:
    conv1d_kernel_size: int = 4
    qkv_proj_blocksize: int = 4
    num_heads: int = 4
    proj_factor: float = 2.0
    embedding_dim: int = -1
    bias: bool = False
    dropout: float = 0.0
    context_length: int = -1
    _num_blocks: int = 1
    _inner_embedding_dim: Optional[int] = None

    def finish(self):
        self._inner_embedding_dim = int(math.ceil(self.proj_factor * self.embedding_dim / 64.0) * 64)


@dataclass
class mLSTMBlockConfig:
    mlstm: mLSTMLayerConfig = field(default_factory=mLSTMLayerConfig)


@dataclass
class sLSTMLayerConfig:
    backend: str = "cuda"
    num_heads: int = 4
    conv1d_kernel_size: int = 4
    bias_init: str = "powerlaw_blockdependent"
    embedding_dim: int = -1
    dropout: float = 0.0
    _num_blocks: int = 1


@dataclass
class FeedForwardConfig:
    proj_factor: float = 1.3
    act_fn: str = "gelu"
    embedding_dim: int = -1
    dropout: float = 0.0
    bias: bool = False
    _num_blocks: int = 1


@dataclass
class sLSTMBlockConfig:
    slstm: sLSTMLayerConfig = field(default_factory=sLSTMLayerConfig)
    feedforward: Optional[FeedForwardConfig] = None


@dataclass
class xLSTMBlockStackConfig:
    mlstm_block: Optional[mLSTMBlockConfig] = None
    slstm_block: Optional[sLSTMBlockConfig] = None
    context_length: int = -1
    num_blocks: int = 1
    embedding_dim: int = 128
    add_post_blocks_norm: bool = True
    bias: bool = False
    dropout: float = 0.0
    slstm_at: List[int] = field(default_factory=list)


# ---- components ------------------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """xlstm.components.ln.LayerNorm: gamma = 1 + weight (residual_weight), bias optional (off in the presets)."""

    def __init__(self, ndim: int = -1, weight: bool = True, bias: bool = False, eps: float = 1e-5,
                 residual_weight: bool = True):
        super().__init__()
        self.ndim, self.eps, self.residual_weight = ndim, eps, residual_weight
        self.weight = nn.Parameter(torch.zeros(ndim)) if weight else None
        self.bias = nn.Parameter(torch.zeros(ndim)) if bias else None

    @property
    def weight_proxy(self):
        return 1.0 + self.weight if self.residual_weight else self.weight

    def forward(self, x):
        return F.layer_norm(x, (self.ndim,), weight=self.weight_proxy, bias=self.bias, eps=self.eps)

    def reset_parameters(self):
        with torch.no_grad():
            self.weight.zero_() if self.residual_weight else self.weight.fill_(1.0)
            if self.bias is not None:
                self.bias.zero_()


class MultiHeadLayerNorm(LayerNorm):
    def forward(self, x):                                  # x [B, NH, S, DH]
        B, NH, S, DH = x.shape
        g = x.transpose(1, 2).reshape(B * S, NH * DH)
        out = F.group_norm(g, num_groups=NH, weight=self.weight_proxy, bias=self.bias, eps=self.eps)
        return out.view(B, S, NH, DH).transpose(1, 2)


class LinearHeadwiseExpand(nn.Module):
    def __init__(self, in_features: int, num_heads: int, bias: bool = False, std: Optional[float] = None):
        super().__init__()
        assert not bias
        self.in_features, self.num_heads = in_features, num_heads
        bs = in_features // num_heads
        self.weight = nn.Parameter(torch.empty(num_heads, bs, bs))
        self._std = std if std is not None else math.sqrt(2.0 / (5.0 * bs))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight, mean=0.0, std=self._std)

    def forward(self, x):
        return O.headwise_linear(x, self.weight)


class CausalConv1d(nn.Module):
    def __init__(self, feature_dim: int, kernel_size: int):
        super().__init__()
        self.conv = nn.Conv1d(feature_dim, feature_dim, kernel_size, padding=kernel_size - 1, groups=feature_dim,
                              bias=True)

    def forward(self, x):
        return O.causal_conv1d_forward(x, self.conv.weight, self.conv.bias)

    def step(self, x, conv_state=None):
        if conv_state is None:
            conv_state = (torch.zeros(x.shape[0], self.conv.kernel_size[0], x.shape[2], dtype=self.conv.weight.dtype,
                                      device=self.conv.weight.device),)
        y, cs = O.conv1d_step(x, conv_state[0], self.conv.weight, self.conv.bias)
        return y, (cs,)

    def reset_parameters(self):
        self.conv.reset_parameters()


class mLSTMCell(nn.Module):
    def __init__(self, inner: int, num_heads: int, context_length: int):
        super().__init__()
        self.inner, self.num_heads = inner, num_heads
        self.igate = nn.Linear(3 * inner, num_heads)
        self.fgate = nn.Linear(3 * inner, num_heads)
        self.outnorm = MultiHeadLayerNorm(ndim=inner, weight=True, bias=False)
        self.register_buffer("causal_mask", torch.tril(torch.ones(max(context_length, 1), max(context_length, 1),
                                                                  dtype=torch.bool)), persistent=False)
        self.reset_parameters()

    def _heads(self, q, k, v):
        B, S, _ = q.shape
        NH = self.num_heads
        g = torch.cat([q, k, v], dim=-1)
        ig = self.igate(g).transpose(-1, -2).unsqueeze(-1)                 # [B,NH,S,1]
        fg = self.fgate(g).transpose(-1, -2).unsqueeze(-1)
        qh, kh, vh = (t.view(B, S, NH, -1).transpose(1, 2) for t in (q, k, v))
        return qh, kh, vh, ig, fg

    def forward(self, q, k, v):
        B, S, _ = q.shape
        qh, kh, vh, ig, fg = self._heads(q, k, v)
        h = O.parallel_stabilized_simple(qh, kh, vh, ig, fg)
        return self.outnorm(h).transpose(1, 2).reshape(B, S, -1)

    def step(self, q, k, v, mlstm_state=None):
        B, S, _ = q.shape
        assert S == 1
        NH, DH = self.num_heads, self.inner // self.num_heads
        qh, kh, vh, ig, fg = self._heads(q, k, v)
        if mlstm_state is None:
            c = torch.zeros(B, NH, DH, DH, device=q.device, dtype=q.dtype)
            n = torch.zeros(B, NH, DH, 1, device=q.device, dtype=q.dtype)
            m = torch.zeros(B, NH, 1, 1, device=q.device, dtype=q.dtype)
        else:
            c, n, m = (t.to(device=q.device, dtype=q.dtype) for t in mlstm_state)
        h, st = O.recurrent_step_stabilized_simple(c, n, m, qh, kh, vh, ig, fg)
        return self.outnorm(h).transpose(1, 2).reshape(B, S, -1), st

    def reset_parameters(self):
        with torch.no_grad():
            self.outnorm.reset_parameters()
            self.fgate.weight.zero_()
            self.fgate.bias.copy_(torch.linspace(3.0, 6.0, self.num_heads))
            self.igate.weight.zero_()
            self.igate.bias.normal_(mean=0.0, std=0.1)


class mLSTMLayer(nn.Module):
    def __init__(self, cfg: mLSTMLayerConfig):
        super().__init__()
        self.config = cfg
        d, inner = cfg.embedding_dim, cfg._inner_embedding_dim
        self.proj_up = nn.Linear(d, 2 * inner, bias=cfg.bias)
        nproj = inner // cfg.qkv_proj_blocksize
        self.q_proj = LinearHeadwiseExpand(inner, nproj)
        self.k_proj = LinearHeadwiseExpand(inner, nproj)
        self.v_proj = LinearHeadwiseExpand(inner, nproj)
        self.conv1d = CausalConv1d(inner, cfg.conv1d_kernel_size)
        self.mlstm_cell = mLSTMCell(inner, cfg.num_heads, cfg.context_length)
        self.learnable_skip = nn.Parameter(torch.ones(inner))
        self.proj_down = nn.Linear(inner, d, bias=cfg.bias)
        self.reset_parameters()

    def _pre(self, x, conv):
        inner = self.config._inner_embedding_dim
        u = self.proj_up(x)
        x_m, z = u[..., :inner], u[..., inner:]
        a = F.silu(conv(x_m))
        return x_m, z, a, self.q_proj(a), self.k_proj(a), self.v_proj(x_m)

    def forward(self, x):
        x_m, z, a, q, k, v = self._pre(x, self.conv1d)
        h = self.mlstm_cell(q, k, v)
        return self.proj_down((h + self.learnable_skip * a) * F.silu(z))

    def step(self, x, mlstm_state=None, conv_state=None):
        box = {}

        def conv(x_m):
            y, box["cs"] = self.conv1d.step(x_m, conv_state)
            return y

        x_m, z, a, q, k, v = self._pre(x, conv)
        h, mlstm_state = self.mlstm_cell.step(q, k, v, mlstm_state)
        y = self.proj_down((h + self.learnable_skip * a) * F.silu(z))
        return y, {"mlstm_state": mlstm_state, "conv_state": box["cs"]}

    def reset_parameters(self):
        d, nb = self.config.embedding_dim, self.config._num_blocks
        with torch.no_grad():
            small = math.sqrt(2.0 / (5.0 * d))
            self.proj_up.weight.normal_(0.0, small)
            self.proj_down.weight.normal_(0.0, 2.0 / (nb * math.sqrt(d)))
            for p in (self.q_proj, self.k_proj, self.v_proj):
                p.weight.normal_(0.0, small)
            self.learnable_skip.fill_(1.0)
            self.conv1d.reset_parameters()
            self.mlstm_cell.reset_parameters()


class sLSTMCell_cuda(nn.Module):
    """Parameter holder with xlstm's names (`_recurrent_kernel_` [NH, DH, 4, DH], `_bias_` [NH, 4, DH]); the
    reference subclasses it only to manage pickling of the CUDA extension (decision_xlstm.py:40-101)."""

    def __init__(self, config, skip_backend_init: bool = False):
        super().__init__()
        self.config = config
        NH, DH = config.num_heads, config.embedding_dim // config.num_heads
        self._recurrent_kernel_ = nn.Parameter(torch.zeros(NH, DH, 4, DH))
        self._bias_ = nn.Parameter(torch.zeros(NH, 4, DH))
        self.func = None
        self.reset_parameters()

    def reset_parameters(self):
        NH, DH = self._bias_.shape[0], self._bias_.shape[2]
        with torch.no_grad():
            self._recurrent_kernel_.normal_(0.0, 0.5 / math.sqrt(DH))
            self._bias_.zero_()
            self._bias_[:, 1, :] = torch.linspace(3.0, 6.0, DH)


class sLSTMLayer(nn.Module):
    def __init__(self, cfg: sLSTMLayerConfig):
        super().__init__()
        self.config = cfg
        d, NH = cfg.embedding_dim, cfg.num_heads
        self.conv1d = CausalConv1d(d, cfg.conv1d_kernel_size)
        std = math.sqrt(2.0 / (5.0 * d)) * math.sqrt(NH)
        self.fgate = LinearHeadwiseExpand(d, NH, std=std)
        self.igate = LinearHeadwiseExpand(d, NH, std=std)
        self.zgate = LinearHeadwiseExpand(d, NH, std=std)
        self.ogate = LinearHeadwiseExpand(d, NH, std=std)
        self.slstm_cell = sLSTMCell_cuda(cfg)
        self.group_norm = MultiHeadLayerNorm(ndim=d, weight=True, bias=False)

    def step(self, x, slstm_state=None, conv_state=None):
        B, _, d = x.shape
        NH = self.config.num_heads
        DH = d // NH
        xc, cs = self.conv1d.step(x, conv_state)
        xc = F.silu(xc)
        # xlstm v1.0.x sLSTMLayer.forward: i, f, z, o = fgate(x_conv), igate(x_conv), zgate(x), ogate(x)
        wx = torch.cat([self.fgate(xc), self.igate(xc), self.zgate(x), self.ogate(x)], dim=-1).view(B, 4, NH, DH)
        st = torch.zeros(4, B, d) if slstm_state is None else slstm_state
        ry = torch.einsum("bhd,hdgo->bgho", st[0].view(B, NH, DH), self.slstm_cell._recurrent_kernel_)
        raw = (wx + ry + self.slstm_cell._bias_.permute(1, 0, 2).unsqueeze(0)).reshape(B, 4, d)
        st = O.slstm_pointwise(raw, st)
        out = self.group_norm(st[0].view(B, NH, 1, DH))
        return out.transpose(1, 2).reshape(B, 1, d), {"slstm_state": st, "conv_state": cs}

    def forward(self, x):
        outs, st = [], {}
        for t in range(x.shape[1]):
            y, st = self.step(x[:, t:t + 1], **st)
            outs.append(y)
        return torch.cat(outs, dim=1)

    def reset_parameters(self):
        self.slstm_cell.reset_parameters()
        self.group_norm.reset_parameters()
        self.conv1d.reset_parameters()
        for g in (self.fgate, self.igate, self.zgate, self.ogate):
            g.reset_parameters()


class GatedFeedForward(nn.Module):
    def __init__(self, cfg: FeedForwardConfig):
        super().__init__()
        self.ff = int(math.ceil(cfg.proj_factor * cfg.embedding_dim / 64.0) * 64)
        self.proj_up = nn.Linear(cfg.embedding_dim, 2 * self.ff, bias=cfg.bias)
        self.proj_down = nn.Linear(self.ff, cfg.embedding_dim, bias=cfg.bias)

    def forward(self, x):
        u = self.proj_up(x)
        return self.proj_down(F.gelu(u[..., : self.ff]) * u[..., self.ff:])

    def reset_parameters(self):
        d = self.proj_up.in_features
        with torch.no_grad():
            self.proj_up.weight.normal_(0.0, math.sqrt(2.0 / (5.0 * d)))
            self.proj_down.weight.normal_(0.0, math.sqrt(2.0 / (5.0 * d)) * 0.5)


class xLSTMBlock(nn.Module):
    def __init__(self, d: int, layer: nn.Module, ffn: Optional[nn.Module]):
        super().__init__()
        self.xlstm_norm = LayerNorm(ndim=d, weight=True, bias=False)
        self.xlstm = layer
        self.ffn_norm = LayerNorm(ndim=d, weight=True, bias=False) if ffn is not None else None
        self.ffn = ffn

    def forward(self, x):
        x = x + self.xlstm(self.xlstm_norm(x))
        if self.ffn is not None:
            x = x + self.ffn(self.ffn_norm(x))
        return x

    def step(self, x, **kwargs):
        y, st = self.xlstm.step(self.xlstm_norm(x), **kwargs)
        x = x + y
        if self.ffn is not None:
            x = x + self.ffn(self.ffn_norm(x))
        return x, st

    def reset_parameters(self):
        self.xlstm_norm.reset_parameters()
        self.xlstm.reset_parameters()
        if self.ffn is not None:
            self.ffn_norm.reset_parameters()
            self.ffn.reset_parameters()


class xLSTMBlockStack(nn.Module):
    def __init__(self, config: xLSTMBlockStackConfig):
        super().__init__()
        self.config = config
        d, L = config.embedding_dim, config.num_blocks
        blocks = []
        for i in range(L):
            if i in config.slstm_at:
                sc = dataclasses.replace(config.slstm_block.slstm, embedding_dim=d, _num_blocks=L)
                fc = config.slstm_block.feedforward
                ffn = GatedFeedForward(dataclasses.replace(fc, embedding_dim=d, _num_blocks=L)) if fc else None
                blocks.append(xLSTMBlock(d, sLSTMLayer(sc), ffn))
            else:
                mc = dataclasses.replace(config.mlstm_block.mlstm, embedding_dim=d, _num_blocks=L, bias=config.bias,
                                         context_length=config.context_length)
                mc.finish()
                blocks.append(xLSTMBlock(d, mLSTMLayer(mc), None))
        self.blocks = nn.ModuleList(blocks)
        self.post_blocks_norm = LayerNorm(ndim=d) if config.add_post_blocks_norm else nn.Identity()

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return self.post_blocks_norm(x)

    def step(self, x, state=None):
        if state is None:
            state = {}
        for i, b in enumerate(self.blocks):
            x, state[f"block_{i}"] = b.step(x, **state.get(f"block_{i}", {}))
        return self.post_blocks_norm(x), state

    def reset_parameters(self):
        for b in self.blocks:
            b.reset_parameters()
        if isinstance(self.post_blocks_norm, LayerNorm):
            self.post_blocks_norm.reset_parameters()
