"""Batched rollout loop: the B-env restatement of `custom_evaluate_policy`
(src/callbacks/evaluation.py:14-271, which asserts num_envs == 1 at :80).

Per env, the reference's bookkeeping in inference-cache mode reduces to:
  * feed (state_t, rtg_t, reward placeholder 0) for the newest timestep only (:131-139, history pruned :172-177)
  * rtg_{t+1} = rtg_t - r_t / reward_scale                                       (:152-168)
  * on done: rtg <- target return, state history <- reset obs, past_key_values <- None (:238-251)
Here `past_key_values = None` for one env of the batch is a per-env reset mask handed to the library.

Multi-GPU: envs are sharded like the reference shards eval envs, `idx % world_size == rank`
(src/callbacks/custom_eval_callback.py:385,445); results go to every rank with ONE all_gather of the
action tokens / returns (the reference pickles dicts through gather_object, src/utils/misc.py:159-191).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import XLSTMEngine


def shard_env_ids(n_envs: int, rank: int, world_size: int) -> List[int]:
    """custom_eval_callback.py:385: env idx handled by rank idx % world_size."""
    return [i for i in range(n_envs) if i % world_size == rank]


def gather_env_results(local: torch.Tensor, n_envs: int, rank: int, world_size: int, group=None) -> torch.Tensor:
    """All-gather per-env rows (action tokens [B_local, A] or returns [B_local]) and put them back in global env
    order. One collective; works with NCCL (cuda tensors) and gloo (cpu tensors). Shards may be ragged."""
    import torch.distributed as dist
    if world_size == 1:
        return local
    max_local = (n_envs + world_size - 1) // world_size
    pad_shape = (max_local, *local.shape[1:])
    padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world_size, *pad_shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered.view(-1, *local.shape[1:]), padded, group=group)
    out = torch.zeros((n_envs, *local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world_size):
        ids = shard_env_ids(n_envs, r, world_size)
        out[ids] = gathered[r, : len(ids)]
    return out


class OverlappedTokenGather:
    """Per-step all-gather of the action tokens, off the critical path.

    Step t's tokens are snapshotted into one of two staging buffers on the compute stream (a 2 KB copy) and
    all-gathered on a side stream while step t+1 computes; envs never wait for other ranks' actions, so the
    only ordering needed is "staging buffer k is not overwritten before its gather finished" (2 steps apart).
    Requires equal shards (B_local * world == n_envs); results are in rank-major order [world, B_local, A] and
    `global_view()` puts them back in env order (env i lives on rank i % world).
    """

    def __init__(self, B_local: int, act_dim: int, world_size: int, device, group=None):
        self.world, self.group = world_size, group
        self.side = torch.cuda.Stream(device=device)
        self.stage = [torch.zeros(B_local, act_dim, dtype=torch.int32, device=device) for _ in range(2)]
        self.out = [torch.zeros(world_size, B_local, act_dim, dtype=torch.int32, device=device) for _ in range(2)]
        self.done = [None, None]
        self.k = 0

    def submit(self, tokens: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        k = self.k
        cur = torch.cuda.current_stream(tokens.device)
        if self.done[k] is not None:
            cur.wait_event(self.done[k])                 # gather of step t-2 read this staging buffer
        self.stage[k].copy_(tokens, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.out[k].view(-1, tokens.shape[-1]), self.stage[k], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        self.k ^= 1
        return self.out[k]

    def finish(self):
        for ev in self.done:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def global_view(self, gathered: torch.Tensor) -> torch.Tensor:
        # [world, B_local, A] -> [n_envs, A] with env = b * world + r
        return gathered.permute(1, 0, 2).reshape(-1, gathered.shape[-1])


class BatchedRollout:
    """Drives B envs of a `SyntheticEnvBatch`-like object (reset() -> obs[B,204]; step(a) -> obs, reward, done)
    through the CUDA policy, one library call per env step with pinned host buffers."""

    def __init__(self, engine: XLSTMEngine, envs, mode: int = L.XL_MODE_FUSED, use_graph: bool = True):
        self.engine, self.envs = engine, envs
        self.B = envs.B
        cfg = engine.cfg
        self.mode = mode
        self.flags = L.XL_FLAG_GRAPH if use_graph else 0
        self.state = engine.new_state(self.B)
        pin = dict(pin_memory=True)
        self.h_states = torch.zeros(self.B, cfg.state_dim, dtype=torch.float32, **pin)
        self.h_rtg = torch.zeros(self.B, dtype=torch.float32, **pin)
        self.h_tokens = torch.zeros(self.B, cfg.act_dim, dtype=torch.int32, **pin)
        self.h_actions = torch.zeros(self.B, cfg.act_dim, dtype=torch.float32, **pin)
        self.rtg0 = np.asarray(envs.rtg0, dtype=np.float32)
        self.scale = np.asarray(envs.reward_scale, dtype=np.float32)

    def run(self, n_steps: int, record: bool = False) -> Dict[str, object]:
        envs, B = self.envs, self.B
        obs = envs.reset()
        rtg = self.rtg0.copy()
        self.engine.reset(self.state)
        ep_ret = np.zeros(B, dtype=np.float64)
        returns: List[List[float]] = [[] for _ in range(B)]
        toks = [] if record else None
        acts = [] if record else None
        t0 = time.perf_counter()
        for _ in range(n_steps):
            self.h_states.numpy()[:] = obs
            self.h_rtg.numpy()[:] = rtg
            self.engine.policy_step_host(self.state, self.h_states, self.h_rtg, self.h_tokens, self.h_actions,
                                         mode=self.mode, flags=self.flags)
            actions = self.h_actions.numpy()
            if record:
                toks.append(self.h_tokens.numpy().copy())
                acts.append(actions.copy())
            # env_act_dim truncation (discrete_decision_transformer_sb3.py:70-71) is per env
            obs, reward, done = envs.step(actions)
            ep_ret += reward
            rtg = (rtg - reward / self.scale).astype(np.float32)
            if done.any():
                for i in np.nonzero(done)[0]:
                    returns[i].append(float(ep_ret[i]))
                    ep_ret[i] = 0.0
                rtg[done] = self.rtg0[done]
                self.engine.reset(self.state, torch.from_numpy(done.astype(np.uint8)))
        dt = time.perf_counter() - t0
        out: Dict[str, object] = {
            "env_steps": n_steps * B, "seconds": dt,
            # the reference's metric names (custom_eval_callback.py:468-475)
            "steps_per_second": n_steps / dt, "total_steps_per_second": n_steps * B / dt,
            "episode_returns": returns,
        }
        if record:
            out["tokens"] = np.stack(toks)
            out["actions"] = np.stack(acts)
        return out
