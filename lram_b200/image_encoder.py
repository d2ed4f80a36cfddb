"""Image front-end of the policy: `embed_image` = IMPALA CNN.

Reference: `ImpalaCNN` / `ImpalaCNNBlock` / `ImpalaCNNResidual` (src/algos/models/image_encoders.py:10-125), built by
`make_image_encoder` for `image_shape=[3,64,64]` (multi_domain_discrete_dt_model.py:38-46,
configs/agent_params/model_kwargs/multi_domain.yaml) and applied to `states.float() / 255`
(online_decision_transformer_model.py:522-526) one timestep at a time on the rollout path
(discrete_decision_transformer_model.py:187-203).

Per SURVEY.md §8 a12 this stage stays in PyTorch / cuDNN (it is not on the HBM-bound recurrent path): three
stages of conv3x3 -> maxpool(3, stride 2, pad 1) -> 2 residual units (relu -> conv3x3 -> relu -> conv3x3, skip),
channels 16/32/32 x model_size, then ReLU, flatten, Linear(n_flat -> d), ReLU. Parameter names equal the
reference's, so `embed_image.*` entries of an LRAM checkpoint load with `load_state_dict`. The resulting state
embeddings [B, d] enter the CUDA path through `xl_policy_step(..., XL_FLAG_STATE_EMBEDS)`.
(Spectral-norm and modulation-vector variants of the reference class are training options, unused by the
multi_domain configuration, and are not reproduced.)
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Residual(nn.Module):
    def __init__(self, depth: int):
        super().__init__()
        self.conv_0 = nn.Conv2d(depth, depth, 3, 1, 1)
        self.conv_1 = nn.Conv2d(depth, depth, 3, 1, 1)

    def forward(self, x):
        y = self.conv_0(F.relu(x))
        y = self.conv_1(F.relu(y))
        return x + y


class _Stage(nn.Module):
    def __init__(self, depth_in: int, depth_out: int):
        super().__init__()
        self.conv = nn.Conv2d(depth_in, depth_out, 3, 1, 1)
        self.residual_0 = _Residual(depth_out)
        self.residual_1 = _Residual(depth_out)

    def forward(self, x):
        x = F.max_pool2d(self.conv(x), 3, 2, padding=1)
        return self.residual_1(self.residual_0(x))


class ImpalaCNN(nn.Module):
    def __init__(self, image_shape: Sequence[int] = (3, 64, 64), features_dim: int = 512, model_size: int = 1,
                 out_relu: bool = True):
        super().__init__()
        c, hgt, wid = image_shape
        self.image_shape = tuple(image_shape)
        self.cnn = nn.ModuleList([_Stage(c, 16 * model_size), _Stage(16 * model_size, 32 * model_size),
                                  _Stage(32 * model_size, 32 * model_size)])
        for _ in range(3):                                   # maxpool(3, 2, pad 1): n -> floor((n - 1) / 2) + 1
            hgt, wid = (hgt - 1) // 2 + 1, (wid - 1) // 2 + 1
        self.n_flatten = 32 * model_size * hgt * wid
        self.linear = nn.Sequential(nn.Linear(self.n_flatten, features_dim), nn.ReLU() if out_relu else nn.Identity())

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, C, H, W] float, ALREADY scaled to 0..1: the caller divides raw frames by 255 exactly once, as the
        reference's `embed_inputs` does before `embed_image` (online_decision_transformer_model.py:522-526)."""
        if not x.is_floating_point():
            raise TypeError("ImpalaCNN takes float frames already scaled by 1/255 (see scale_frames)")
        for stage in self.cnn:
            x = stage(x)
        return self.linear(torch.flatten(F.relu(x), 1))


def scale_frames(frames: torch.Tensor) -> torch.Tensor:
    """`states.float() / 255.0` (online_decision_transformer_model.py:523-525), unconditionally: image observations
    reach the policy as raw 0..255 values whether their dtype is uint8 or — after `pad_inputs` returned
    `states.float()` (src/algos/decision_xlstm.py:25) — float."""
    return frames.float() / 255.0


def make_impala_state_dict(features_dim: int, image_shape: Sequence[int] = (3, 64, 64), seed: int = 0,
                           prefix: str = "embed_image.") -> Dict[str, torch.Tensor]:
    """Seeded default-init weights with the reference's key names (for synthetic policies and tests)."""
    g = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng():
        torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
        net = ImpalaCNN(image_shape, features_dim)
    return {prefix + k: v.detach().clone() for k, v in net.state_dict().items()}


def split_image_weights(sd: Dict[str, torch.Tensor], prefix: str = "embed_image.") -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
