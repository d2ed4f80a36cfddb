"""Shapes of the xLSTM policy on the recurrent-inference path.

Mirrors the keys the reference reads from its Hydra tree for this path:
  * `configs/agent_params/huggingface/xlstm_{medium,mediumplus,large,huge}.yaml:1-26`
    (hidden_size, n_layer, n_head, xlstm_config.mlstm_block.mlstm.{conv1d_kernel_size,
    qkv_proj_blocksize,num_heads}, num_blocks, embedding_dim)
  * `configs/agent_params/model_kwargs/multi_domain.yaml:1-11`
    (action_channels=256, discrete_actions=18, state_dim=204, shared_a_head, reward_condition,
    action_condition=False, use_time_embds=False)
  * `configs/agent_params/replay_buffer_kwargs/multi_domain_mtdmccs.yaml:2-3`
    (max_state_dim=204, max_act_dim=8)
and the defaults of the third-party `xlstm` v1.0.x mLSTM layer the reference builds at
`src/algos/models/decision_xlstm.py:130-133` (proj_factor 2.0 rounded up to 64, bias False,
post-blocks norm, LN eps 1e-5, cell eps 1e-6).
"""
from __future__ import annotations

import dataclasses
import math


@dataclasses.dataclass(frozen=True)
class XLSTMPolicyConfig:
    # encoder (xlstm_config)
    embedding_dim: int = 512          # d
    num_blocks: int = 8               # L
    num_heads: int = 4                # NH
    conv1d_kernel_size: int = 4       # KS
    qkv_proj_blocksize: int = 4       # bs
    proj_factor: float = 2.0
    ln_eps: float = 1e-5              # xlstm LayerNorm / MultiHeadLayerNorm eps
    cell_eps: float = 1e-6            # recurrent_step_stabilized_simple eps
    # xLSTM[a:b] stacks (`xlstm_ms_mediumplus.yaml:14-27`): blocks listed here are sLSTM blocks
    # (sLSTM layer + gated feed-forward), every other block is an mLSTM block without feed-forward
    slstm_at: tuple = ()
    ffn_proj_factor: float = 1.3      # slstm_block.feedforward.proj_factor (act_fn gelu)
    # LRAM policy around it (multi_domain model kwargs)
    state_dim: int = 204              # max_state_dim, embed_state = Linear(204, d)
    act_dim: int = 8                  # max_act_dim (config.act_dim)
    action_channels: int = 256        # tokenizer vocab
    discrete_actions: int = 18        # tokenizer shift / discrete head width
    embed_ln_eps: float = 1e-5        # HF nn.LayerNorm(hidden_size) default
    tokens_per_step: int = 3          # (s, rtg, r)  discrete_decision_transformer_model.py:266-275
    action_token_pos: int = 1         # tok_to_pred_pos["a"] = s_dim = 1  (:272)
    name: str = "custom"

    @property
    def d(self) -> int:
        return self.embedding_dim

    @property
    def inner(self) -> int:
        # xlstm mLSTMLayerConfig: round proj_factor*d up to a multiple of 64
        return int(math.ceil(self.proj_factor * self.embedding_dim / 64.0) * 64)

    @property
    def head_dim(self) -> int:
        return self.inner // self.num_heads

    @property
    def ffn_dim(self) -> int:
        # xlstm UpProjConfigMixin: round proj_factor*d up to a multiple of 64
        return int(math.ceil(self.ffn_proj_factor * self.embedding_dim / 64.0) * 64)

    @property
    def slstm_head_dim(self) -> int:
        return self.embedding_dim // self.num_heads

    def is_slstm(self, i: int) -> bool:
        return i in self.slstm_at

    @property
    def slstm_mask(self) -> int:
        m = 0
        for i in self.slstm_at:
            m |= 1 << int(i)
        return m

    @property
    def num_actions(self) -> int:
        # multi_domain_discrete_dt_model.py:31
        return self.discrete_actions + self.action_channels

    @property
    def head_out(self) -> int:
        # shared action head: num_actions * act_dim   (multi_domain_discrete_dt_model.py:67-71)
        return self.num_actions * self.act_dim

    def validate(self) -> None:
        assert self.inner % self.num_heads == 0
        assert self.inner % self.qkv_proj_blocksize == 0
        assert self.head_dim % 4 == 0, "head_dim must be a multiple of 4 (128-bit state accesses)"
        assert self.embedding_dim % 4 == 0
        assert 1 <= self.conv1d_kernel_size <= 8
        assert all(0 <= i < self.num_blocks and i < 64 for i in self.slstm_at), "slstm_at out of range"
        if self.slstm_at:
            assert self.embedding_dim % self.num_heads == 0

    # ---- bytes / flops used by bench.py and DESIGN.md -----------------------
    def state_bytes_per_env_layer(self, i: int = -1) -> int:
        """Resident fp32 state of one env in one block: C + n + m + conv window (KS rows); an sLSTM block
        holds (y, c, n, m) [4, d] + its conv window [KS, d]."""
        if i >= 0 and self.is_slstm(i):
            return 4 * (4 * self.d + self.conv1d_kernel_size * self.d)
        nh, dh = self.num_heads, self.head_dim
        return 4 * (nh * dh * dh + nh * dh + nh + self.conv1d_kernel_size * self.inner)

    def state_bytes_per_env(self) -> int:
        return sum(self.state_bytes_per_env_layer(i) for i in range(self.num_blocks))

    def algorithmic_bytes_per_env_layer_tokenstep(self) -> int:
        """8*NH*DH^2 + 8*NH*DH + 8*NH + 16*inner (C,n,m R+W; conv 3R+1W)."""
        nh, dh = self.num_heads, self.head_dim
        return 8 * nh * dh * dh + 8 * nh * dh + 8 * nh + 16 * self.inner

    def encoder_params(self) -> int:
        d, inner, nh = self.d, self.inner, self.num_heads
        bs, ks = self.qkv_proj_blocksize, self.conv1d_kernel_size
        per_block = (d + 2 * inner * d + 3 * (inner // bs) * bs * bs + inner * ks + inner
                     + 2 * (nh * 3 * inner + nh) + inner + inner + d * inner)
        ff, dh = self.ffn_dim, (d // nh if self.slstm_at else 0)
        # sLSTM block: norm, conv, 4 headwise gates [NH,DH,DH], recurrent kernel [NH,DH,4,DH], bias [NH,4,DH],
        # group norm, ffn norm, gated feed-forward (up [2ff,d], down [d,ff])
        per_slstm = (d + d * ks + d + 4 * nh * dh * dh + nh * dh * 4 * dh + nh * 4 * dh + d + d + 2 * ff * d + d * ff)
        ns = len(self.slstm_at)
        return (self.num_blocks - ns) * per_block + ns * per_slstm + d


_PRESETS = {
    # README.md:183,198,213,228 / xlstm_{medium,mediumplus,large,huge}.yaml
    "16M": dict(embedding_dim=512, num_blocks=8),
    "48M": dict(embedding_dim=768, num_blocks=12),
    "110M": dict(embedding_dim=1024, num_blocks=16),
    "206M": dict(embedding_dim=1280, num_blocks=20),
    # xLSTM[7:1]-style stack of `xlstm_ms_mediumplus.yaml:27` (README.md:188-190): sLSTM block at position 1
    "48M-ms": dict(embedding_dim=768, num_blocks=12, slstm_at=(1,)),
    "toy-ms": dict(embedding_dim=64, num_blocks=3, slstm_at=(1,)),
    "toy128-ms": dict(embedding_dim=128, num_blocks=4, slstm_at=(0, 2)),
    # test-only toy sizes
    "toy": dict(embedding_dim=64, num_blocks=2),
    "toy128": dict(embedding_dim=128, num_blocks=3),
}


def preset(name: str, **overrides) -> XLSTMPolicyConfig:
    if name not in _PRESETS:
        raise KeyError(f"unknown preset {name!r}; have {sorted(_PRESETS)}")
    kw = dict(_PRESETS[name])
    kw.update(overrides)
    cfg = XLSTMPolicyConfig(name=name, **kw)
    cfg.validate()
    return cfg
