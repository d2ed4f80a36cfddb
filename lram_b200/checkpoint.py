"""Checkpoint -> weight blob for the CUDA path.

The reference stores agents in the stable-baselines3 zip layout (`save_to_zip_file_fixed`,
src/algos/agent_utils.py:165-202): a zip archive holding

    data                    JSON of the non-torch attributes (ignored here)
    pytorch_variables.pth   torch.save({"state_mean": ..., "state_std": ...})   (decision_transformer_sb3.py:1115-1118)
    policy.pth              torch.save(policy.state_dict())
    optimizer.pth           (ignored)
    system_info.txt

and restores only the policy weights through `load_model_weights` (decision_transformer_sb3.py:1120-1184): strip a
leading "module." (DDP), drop excluded heads, strip "_orig_mod." (torch.compile), `load_state_dict(strict=False)`,
pick up state_mean / state_std. This module restates that loader without stable-baselines3 (not installed) and
derives the `XLSTMPolicyConfig` from the tensor shapes, so a released LRAM checkpoint can be bound to the engine:

    sd, variables = load_policy_zip("model.zip")
    cfg = infer_config(sd)
    engine = XLSTMEngine(cfg, sd, max_batch=B)

`save_policy_zip` writes the same layout (used by the tests and by tools that export synthetic weights).
A bare `torch.save(state_dict)` file (.pt/.pth) is accepted too.
"""
from __future__ import annotations

import io
import json
import re
import zipfile
from typing import Dict, Optional, Tuple

import torch

from .config import XLSTMPolicyConfig

_HEAD_BASES = ("action_net", "action_pred", "mu", "log_std")


def _excluded(load_action_head: bool, load_state_head: bool):
    ex = []
    if not load_action_head:                      # decision_transformer_sb3.py:1124-1129
        for name in _HEAD_BASES:
            ex += [f"{name}.weight", f"{name}.bias", f"{name}.0.weight", f"{name}.0.bias", f"{name}.1.weight",
                   f"{name}.1.bias"]
    if not load_state_head:                       # :1130-1131
        ex += ["predict_state.weight", "predict_state.bias"]
    return set(ex)


def fix_policy_keys(policy_dict: Dict[str, torch.Tensor], load_action_head: bool = True,
                    load_state_head: bool = False) -> Dict[str, torch.Tensor]:
    """Key fix-ups of `load_model_weights` for an uncompiled target policy (decision_transformer_sb3.py:1137-1158):
    first "module." occurrence removed, excluded heads dropped, first "_orig_mod." occurrence removed, legacy
    mu/log_std names moved under `.0.`."""
    ex = _excluded(load_action_head, load_state_head)
    out = {}
    for k, v in policy_dict.items():
        k = k.replace("module.", "", 1).replace("_orig_mod.", "", 1)
        if k in ex:       # (the reference tests the exclusion list before stripping "_orig_mod."; none of the
            continue      #  excludable heads is read by this path, so the order does not matter here)
        out[k] = v
    if "mu.weight" in out:
        for base in ("mu", "log_std"):
            out[f"{base}.0.weight"] = out.pop(f"{base}.weight")
            out[f"{base}.0.bias"] = out.pop(f"{base}.bias")
    return out


def _torch_load(fh) -> dict:
    data = fh.read() if hasattr(fh, "read") else fh
    return torch.load(io.BytesIO(data), map_location="cpu", weights_only=True)


def load_policy_zip(path: str, load_action_head: bool = True, load_state_head: bool = False
                    ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """-> (policy state_dict with reference key names, variables {"state_mean", "state_std"} when stored)."""
    variables: Dict[str, torch.Tensor] = {}
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            names = set(z.namelist())
            if "policy.pth" in names:                         # SB3 layout
                with z.open("policy.pth") as fh:
                    policy = _torch_load(fh)
                if "pytorch_variables.pth" in names:
                    with z.open("pytorch_variables.pth") as fh:
                        v = _torch_load(fh)
                    if isinstance(v, dict):
                        variables = {k: t for k, t in v.items() if k in ("state_mean", "state_std")}
            else:                                             # torch.save's own zip container
                policy = torch.load(path, map_location="cpu", weights_only=True)
    else:
        policy = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(policy, dict) and "policy" in policy and isinstance(policy["policy"], dict):
        policy = policy["policy"]
    if not isinstance(policy, dict) or not policy:
        raise ValueError(f"{path}: no policy state_dict found")
    return fix_policy_keys(policy, load_action_head, load_state_head), variables


def save_policy_zip(path: str, state_dict: Dict[str, torch.Tensor], state_mean=None, state_std=None,
                    data: Optional[dict] = None, prefix: str = "") -> None:
    """Write the SB3 zip layout of `save_to_zip_file_fixed` (agent_utils.py:165-202). `prefix` (e.g. "module." or
    "_orig_mod.") reproduces DDP / torch.compile checkpoints."""
    def dump(obj) -> bytes:
        buf = io.BytesIO()
        torch.save(obj, buf)
        return buf.getvalue()

    with zipfile.ZipFile(path, mode="w") as z:
        z.writestr("data", json.dumps(data or {}))
        variables = {}
        if state_mean is not None:
            variables["state_mean"] = torch.as_tensor(state_mean)
        if state_std is not None:
            variables["state_std"] = torch.as_tensor(state_std)
        z.writestr("pytorch_variables.pth", dump(variables))
        z.writestr("policy.pth", dump({prefix + k: v.detach().cpu() for k, v in state_dict.items()}))
        z.writestr("system_info.txt", "lram_b200.checkpoint.save_policy_zip\n")


_BLOCK_RE = re.compile(r"^encoder\.layers\.blocks\.(\d+)\.")


def infer_config(sd: Dict[str, torch.Tensor], name: str = "checkpoint") -> XLSTMPolicyConfig:
    """Shapes -> XLSTMPolicyConfig (sLSTM blocks of xLSTM[a:b] stacks are recognised by their
    `xlstm.slstm_cell._recurrent_kernel_` entry). Refuses checkpoints this path does not cover (ln_bias, RMSNorm,
    image-only policies, stacks without any mLSTM block), naming the offending key."""
    blocks = sorted({int(m.group(1)) for k in sd for m in [_BLOCK_RE.match(k)] if m})
    if not blocks or blocks != list(range(len(blocks))):
        raise ValueError("state_dict has no contiguous encoder.layers.blocks.{i}.* entries (not an xLSTM policy?)")
    slstm_at = tuple(i for i in blocks if f"encoder.layers.blocks.{i}.xlstm.slstm_cell._recurrent_kernel_" in sd)
    mlstm = [i for i in blocks if i not in slstm_at]
    if not mlstm:
        raise NotImplementedError("a stack of sLSTM blocks only is not on this path (no mLSTM block found)")
    for k in sd:
        m = _BLOCK_RE.match(k)
        if m and ".ffn." in k and int(m.group(1)) not in slstm_at:
            raise NotImplementedError(f"{k}: feed-forward sub-layer on an mLSTM block is not on this path")
        if k.endswith("xlstm_norm.bias") or k.endswith("outnorm.bias") or k.endswith("post_blocks_norm.bias"):
            raise NotImplementedError(f"{k}: ln_bias=True variants are not supported")
    b0 = f"encoder.layers.blocks.{mlstm[0]}.xlstm."
    up = sd[b0 + "proj_up.weight"]
    d = up.shape[1]
    inner = up.shape[0] // 2
    ig = sd[b0 + "mlstm_cell.igate.weight"]
    nh = ig.shape[0]
    if ig.shape[1] != 3 * inner:
        raise ValueError(f"igate.weight {tuple(ig.shape)} does not match inner={inner}")
    conv = sd[b0 + "conv1d.conv.weight"]
    ks = conv.shape[-1]
    qw = sd[b0 + "q_proj.weight"]
    bs = qw.shape[-1]
    if "embed_state.weight" not in sd:
        raise NotImplementedError("no embed_state.weight: image-only policies (ImpalaCNN) enter through "
                                  "state embeddings, see XLSTMEngine.policy_step_embedded")
    state_dim = sd["embed_state.weight"].shape[1]
    head = sd["action_net.0.weight"] if "action_net.0.weight" in sd else None
    cfg_kw = dict(embedding_dim=d, num_blocks=len(blocks), num_heads=nh, conv1d_kernel_size=ks,
                  qkv_proj_blocksize=bs, proj_factor=inner / d, state_dim=state_dim, name=name)
    if slstm_at:
        s0 = f"encoder.layers.blocks.{slstm_at[0]}."
        rk = sd[s0 + "xlstm.slstm_cell._recurrent_kernel_"]
        if tuple(rk.shape) != (nh, d // nh, 4, d // nh):
            raise ValueError(f"_recurrent_kernel_ {tuple(rk.shape)} is not [NH, d/NH, 4, d/NH] for d={d}, NH={nh}")
        ff = sd[s0 + "ffn.proj_down.weight"].shape[1]
        cfg_kw.update(slstm_at=slstm_at, ffn_proj_factor=ff / d)
        if XLSTMPolicyConfig(**cfg_kw).ffn_dim != ff:
            raise ValueError(f"ffn dim {ff} is not proj_factor*d rounded up to 64 (d={d})")
    cfg = XLSTMPolicyConfig(**cfg_kw)
    if cfg.inner != inner:
        raise ValueError(f"inner={inner} is not proj_factor*d rounded up to 64 (d={d})")
    if head is not None and head.shape[0] != cfg.head_out:
        # shared head = (discrete_actions + action_channels) * act_dim; keep the defaults 18 + 256 and solve act_dim
        na = cfg.num_actions
        if head.shape[0] % na:
            raise ValueError(f"action_net.0.weight has {head.shape[0]} rows, not a multiple of {na}")
        cfg = XLSTMPolicyConfig(**{**cfg_kw, "act_dim": head.shape[0] // na})
    cfg.validate()
    return cfg
