"""Deterministic synthetic weights and observation streams (no network, no checkpoints).

Weights are produced as a flat ``state_dict`` with the reference's own parameter names, i.e. what
``MultiDomainDiscreteDecisionXLSTMModel.state_dict()`` holds for the hot path
(`src/algos/models/decision_xlstm.py:175-234`, `multi_domain_discrete_dt_model.py:12-81`,
`online_decision_transformer_model.py:92-94`; xlstm block names per SURVEY.md Appendix A), so that the
same dict can be fed to the CPU oracle and to the CUDA path, and a real LRAM checkpoint can take its place.

Init follows the xlstm v1.0.x ``reset_parameters`` recipe with the gate weights
perturbed so that the gates are input dependent (otherwise igate/fgate would be constants and a broken
gate GEMV would go unnoticed).

Streams follow BASELINE.md §2b: Meta-World-like / DMControl-like / Composuite-like / Mimicgen-like
state vectors, U(-1,1) in the task's slots and zero elsewhere, zero padded to 204 exactly as
`src/algos/decision_xlstm.py:16-24` pads them; reward 1 per step like `src/envs/dummy_env_utils.py:28`.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import XLSTMPolicyConfig

# tensors the CUDA path keeps in bf16 (large GEMM operands). Everything else stays fp32.
BF16_WEIGHT_SUFFIXES = (
    "xlstm.proj_up.weight",
    "xlstm.proj_down.weight",
    "embed_state.weight",
    "action_net.0.weight",
    "ffn.proj_up.weight",        # sLSTM block's gated feed-forward
    "ffn.proj_down.weight",
)


def is_bf16_weight(name: str) -> bool:
    return name.endswith(BF16_WEIGHT_SUFFIXES)


def round_weights_to_bf16_(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Round the GEMM matrices to bf16-representable fp32 values, in place.

    north_star: "bf16 weights, fp32 state". Oracle and CUDA path must see the *same* numbers, so the
    rounding is applied once to the shared state_dict; the oracle then computes in fp32 on them.
    """
    for k, v in sd.items():
        if is_bf16_weight(k):
            sd[k] = v.to(torch.bfloat16).to(torch.float32)
    return sd


def make_state_dict(cfg: XLSTMPolicyConfig, seed: int = 0, gate_std: float = 0.02,
                    round_bf16: bool = True) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    d, inner, nh, L = cfg.d, cfg.inner, cfg.num_heads, cfg.num_blocks
    bs, ks = cfg.qkv_proj_blocksize, cfg.conv1d_kernel_size

    def normal(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean

    def uniform(*shape, bound=1.0):
        return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound

    sd: Dict[str, torch.Tensor] = {}
    init_range = 0.02  # HF DecisionTransformerConfig.initializer_range
    # --- LRAM embeddings (online_decision_transformer_model.py:92-94, multi_domain...:47-49)
    sd["embed_state.weight"] = normal(d, cfg.state_dim, std=init_range)
    sd["embed_state.bias"] = normal(d, std=init_range)
    sd["embed_return.weight"] = normal(d, 1, std=init_range)
    sd["embed_return.bias"] = normal(d, std=init_range)
    sd["embed_rewards.weight"] = normal(d, 1, std=init_range)
    sd["embed_rewards.bias"] = normal(d, std=init_range)
    sd["embed_ln.weight"] = 1.0 + normal(d, std=0.05)
    sd["embed_ln.bias"] = normal(d, std=init_range)
    # universal action embedding (computed but never fed: action_condition=False)
    sd["embed_action_disc.weight"] = normal(cfg.num_actions + 1, d, std=init_range)
    # --- action head (make_head with n_layer_head=1 -> nn.Sequential(Linear))
    sd["action_net.0.weight"] = normal(cfg.head_out, d, std=0.05)
    sd["action_net.0.bias"] = normal(cfg.head_out, std=init_range)
    # --- xLSTM block stack
    small = math.sqrt(2.0 / (5.0 * d))
    wang = 2.0 / (L * math.sqrt(d))
    for i in range(L):
        p = f"encoder.layers.blocks.{i}."
        if cfg.is_slstm(i):
            # sLSTM block (xlstm v1.0.x sLSTMLayer + GatedFeedForward; names as in its state_dict)
            dh, ff = d // nh, cfg.ffn_dim
            cb = 1.0 / math.sqrt(ks)
            sd[p + "xlstm_norm.weight"] = normal(d, std=0.05)
            sd[p + "xlstm.conv1d.conv.weight"] = uniform(d, 1, ks, bound=cb)
            sd[p + "xlstm.conv1d.conv.bias"] = uniform(d, bound=cb)
            for gname in ("fgate", "igate", "zgate", "ogate"):
                sd[p + f"xlstm.{gname}.weight"] = normal(nh, dh, dh, std=small * math.sqrt(nh))
            sd[p + "xlstm.slstm_cell._recurrent_kernel_"] = normal(nh, dh, 4, dh, std=1.0 / math.sqrt(dh) * 0.5)
            bias = normal(nh, 4, dh, std=0.1)
            bias[:, 1, :] += torch.linspace(3.0, 6.0, dh)          # forget-gate bias (powerlaw-like, positive)
            sd[p + "xlstm.slstm_cell._bias_"] = bias
            sd[p + "xlstm.group_norm.weight"] = normal(d, std=0.05)
            sd[p + "ffn_norm.weight"] = normal(d, std=0.05)
            sd[p + "ffn.proj_up.weight"] = normal(2 * ff, d, std=small)
            sd[p + "ffn.proj_down.weight"] = normal(d, ff, std=max(wang, small * 0.5))
            continue
        sd[p + "xlstm_norm.weight"] = normal(d, std=0.05)                 # gamma = 1 + w
        sd[p + "xlstm.proj_up.weight"] = normal(2 * inner, d, std=small)
        sd[p + "xlstm.q_proj.weight"] = normal(inner // bs, bs, bs, std=small * 4)
        sd[p + "xlstm.k_proj.weight"] = normal(inner // bs, bs, bs, std=small * 4)
        sd[p + "xlstm.v_proj.weight"] = normal(inner // bs, bs, bs, std=small * 4)
        cb = 1.0 / math.sqrt(ks)                                           # nn.Conv1d default init
        sd[p + "xlstm.conv1d.conv.weight"] = uniform(inner, 1, ks, bound=cb)
        sd[p + "xlstm.conv1d.conv.bias"] = uniform(inner, bound=cb)
        sd[p + "xlstm.mlstm_cell.igate.weight"] = normal(nh, 3 * inner, std=gate_std)
        sd[p + "xlstm.mlstm_cell.igate.bias"] = normal(nh, std=0.1)
        sd[p + "xlstm.mlstm_cell.fgate.weight"] = normal(nh, 3 * inner, std=gate_std)
        sd[p + "xlstm.mlstm_cell.fgate.bias"] = torch.linspace(3.0, 6.0, nh)
        sd[p + "xlstm.mlstm_cell.outnorm.weight"] = normal(inner, std=0.05)  # gamma = 1 + w
        sd[p + "xlstm.learnable_skip"] = 1.0 + normal(inner, std=0.05)
        sd[p + "xlstm.proj_down.weight"] = normal(d, inner, std=max(wang, small * 0.5))
    sd["encoder.layers.post_blocks_norm.weight"] = normal(d, std=0.05)
    if round_bf16:
        round_weights_to_bf16_(sd)
    return sd


# ------------------------------------------------------------------------------------------------
# observation streams
# ------------------------------------------------------------------------------------------------
DOMAINS = {
    # name: (state slots filled, env act_dim, rtg0, reward_scale)   BASELINE.md §2b / SURVEY §8d
    "metaworld": (slice(0, 39), 4, 100.0, 200.0),     # env_params/mt45.yaml:2-3
    "dmcontrol": (slice(14, 49), 6, 10.0, 100.0),     # dmcontrol_utils.py:44-50 slots, dmcontrol_icl.yaml:2-3
    "composuite": (slice(0, 93), 8, 50.0, 10.0),
    "mimicgen": (slice(0, 168), 7, 10.0, 1.0),        # mimicgen_utils.py:58-68
}
DOMAIN_ORDER = ("metaworld", "dmcontrol", "composuite", "mimicgen")


class SyntheticEnvBatch:
    """B independent DummyEnv-like environments (`src/envs/dummy_env_utils.py:8-36`): observation is a
    fresh U(-1,1) sample every step, reward is 1, episode ends after ``ep_len`` steps.
    Env ``e`` uses ``np.random.default_rng(seed + e)`` so any subset of envs (a rank's shard) sees exactly
    the stream it would see in the full batch.
    """

    def __init__(self, cfg: XLSTMPolicyConfig, env_ids, domains="metaworld", ep_len: int = 1000,
                 seed: int = 1234):
        self.cfg = cfg
        self.env_ids = [int(e) for e in env_ids]
        self.B = len(self.env_ids)
        if isinstance(domains, str):
            domains = [domains] * self.B if domains != "mixed" else \
                [DOMAIN_ORDER[e % 4] for e in self.env_ids]
        self.domains = list(domains)
        self.ep_len = ep_len
        self.rngs = [np.random.default_rng(seed + e) for e in self.env_ids]
        self.t = np.zeros(self.B, dtype=np.int64)
        self.act_dims = np.array([DOMAINS[dn][1] for dn in self.domains], dtype=np.int64)
        self.rtg0 = np.array([DOMAINS[dn][2] for dn in self.domains], dtype=np.float32)
        self.reward_scale = np.array([DOMAINS[dn][3] for dn in self.domains], dtype=np.float32)

    def _obs(self, i: int) -> np.ndarray:
        sl = DOMAINS[self.domains[i]][0]
        o = np.zeros(self.cfg.state_dim, dtype=np.float32)
        n = sl.stop - sl.start
        o[sl] = self.rngs[i].uniform(-1.0, 1.0, size=n).astype(np.float32)
        return o

    def reset(self) -> np.ndarray:
        self.t[:] = 0
        return np.stack([self._obs(i) for i in range(self.B)])

    def step(self, actions: np.ndarray):
        """actions [B, act_dim] are ignored by the dummy dynamics, like the reference's DummyEnv."""
        self.t += 1
        done = self.t >= self.ep_len
        obs = np.stack([self._obs(i) for i in range(self.B)])
        reward = np.ones(self.B, dtype=np.float32)
        self.t[done] = 0
        return obs, reward, done


def make_stream(cfg: XLSTMPolicyConfig, env_ids, n_steps: int, domains="metaworld", seed: int = 1234):
    """Pre-generated stream: states [n_steps, B, 204] and rtg [n_steps, B] (rtg_{t+1} = rtg_t - r/scale,
    `src/callbacks/evaluation.py:166-167`)."""
    envs = SyntheticEnvBatch(cfg, env_ids, domains=domains, ep_len=10 ** 9, seed=seed)
    states = np.zeros((n_steps, envs.B, cfg.state_dim), dtype=np.float32)
    rtg = np.zeros((n_steps, envs.B), dtype=np.float32)
    states[0] = envs.reset()
    cur = envs.rtg0.copy()
    rtg[0] = cur
    for t in range(1, n_steps):
        obs, r, _ = envs.step(None)
        states[t] = obs
        cur = (cur - r / envs.reward_scale).astype(np.float32)
        rtg[t] = cur
    return states, rtg, envs
