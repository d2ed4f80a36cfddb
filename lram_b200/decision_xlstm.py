"""Host-side mirror of the reference's operator / plugin interface for the hot path.

  FusedXLSTMEncoder                      <->  xLSTMEncoder                         src/algos/models/decision_xlstm.py:119-172
  MultiDomainDiscreteDecisionXLSTMModel  <->  same name                            src/algos/models/decision_xlstm.py:281-289
                                              (.forward: online_decision_transformer_model.py:326-390)
  DiscreteDecisionXLSTM (agent)          <->  same name (predict / pad_inputs / get_action_pred)
                                              src/algos/decision_xlstm.py:6-35, decision_transformer_sb3.py:621-667,
                                              discrete_decision_transformer_sb3.py:13-72

Same names, argument meaning and error behaviour as the reference for this path; everything that is not on
the inference-cache rollout path (training, losses, prompts, autoregressive per-dimension decoding) is out of
scope and raises NotImplementedError instead of silently doing something else. Image observations run the
reference's ImpalaCNN in PyTorch/cuDNN (image_encoder.py) and enter the CUDA path as state embeddings.
All arithmetic runs in libxlstm_b200.so; there is no PyTorch fallback.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, Optional

import torch

from . import _lib as L
from .config import XLSTMPolicyConfig
from .engine import StateCache, XLSTMEngine, strip_checkpoint_prefixes
from .image_encoder import ImpalaCNN, split_image_weights
from .tokenizers import make_tokenizer


class _Output(dict):
    """dict with attribute access, like HF ModelOutput: supports out['k'], out.k, out.get('k')."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return None


class FusedXLSTMEncoder:
    """Drop-in for `xLSTMEncoder` on the `use_cache=True` path: `forward(inputs_embeds=..., past_key_values=...,
    use_cache=True)` returns `last_hidden_state` [B, n_tok, d] and the (opaque) `past_key_values`.

    The reference treats `past_key_values` opaquely (only `is None` tests); here it is a
    `StateCache` living on the GPU and updated IN PLACE (the returned object is the one passed in).
    A reference-format dict ({"block_i": {...}}) is accepted too and imported into a fresh cache.
    """

    def __init__(self, engine: XLSTMEngine, mode: int = L.XL_MODE_PER_TOKEN):
        self.engine = engine
        self.config = engine.cfg
        self.mode = mode

    def reset_parameters(self):  # post_init hook of the reference (decision_xlstm.py:211-214); weights are bound
        return None

    def forward(self, input_ids=None, past_key_values=None, attention_mask=None, token_type_ids=None,
                position_ids=None, head_mask=None, inputs_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None):
        if inputs_embeds is None:
            raise ValueError("xLSTM encoder consumes already embedded inputs (inputs_embeds)")
        if not use_cache:
            raise NotImplementedError("parallel (training) form is out of scope of the recurrent-inference path")
        B = inputs_embeds.shape[0]
        if past_key_values is None:
            cache = self.engine.new_state(B)                   # zeros == state None (m starts at 0)
        elif isinstance(past_key_values, StateCache):
            cache = past_key_values
        elif isinstance(past_key_values, dict):
            cache = self.engine.new_state(B)
            cache.load_past_key_values(past_key_values)
        else:
            raise TypeError(f"unsupported past_key_values type {type(past_key_values)}")
        if cache.B != B:
            raise ValueError(f"past_key_values holds {cache.B} envs, inputs have batch {B}")
        x = inputs_embeds.to(self.engine.device, torch.float32)
        if x.shape[1] > 4:
            # a whole context at once: what `chunkwise_step` (decision_xlstm.py:158-159) asks of layers.step
            hs = self.engine.prefill(cache, x)
        else:
            hs = self.engine.encoder_step(cache, x, mode=self.mode)
        return _Output(last_hidden_state=hs, past_key_values=cache, hidden_states=None, attentions=None)

    __call__ = forward


class MultiDomainDiscreteDecisionXLSTMModel:
    """Policy `forward` on the inference-cache path, batched over B envs.

    forward(states[B,T,204], actions[B,T,A], rewards[B,T,1], returns_to_go[B,T,1], timesteps[B,T], attention_mask,
            ..., past_key_values, use_inference_cache=True) -> output with .action_preds [B,1,A], .action_logits,
            .past_key_values, .last_hidden_state — only the LAST timestep is consumed, exactly as
            `compute_inputs` does when the cache is warm (online_decision_transformer_model.py:466-470); on the cold
            first call the reference embeds all T but `handle_inference_cache` trims to the last 3 tokens
            (decision_xlstm.py:225-229), which is the same thing for T == 1 (the rollout's case).
    """

    def __init__(self, config: XLSTMPolicyConfig, state_dict: Dict[str, torch.Tensor], max_batch: int = 1,
                 device=None, mode: int = L.XL_MODE_FUSED, use_graph: bool = False, image_shape=(3, 64, 64),
                 img_is_encoded: bool = False):
        self.config = config
        self.engine = XLSTMEngine(config, state_dict, max_batch=max_batch, device=device)
        # embed_image (ImpalaCNN, multi_domain_discrete_dt_model.py:38-46) exists when the checkpoint carries it;
        # it stays a PyTorch/cuDNN module feeding the CUDA path with state embeddings
        self.embed_image = None
        self.img_is_encoded = img_is_encoded
        img_sd = split_image_weights(strip_checkpoint_prefixes(state_dict))
        if img_sd:
            self.embed_image = ImpalaCNN(image_shape, config.d).to(self.engine.device).eval()
            self.embed_image.load_state_dict(img_sd)
        self.encoder = FusedXLSTMEncoder(self.engine, mode=mode)
        self.mode = mode
        self.use_graph = use_graph
        self.is_discrete = True            # policy built on a Discrete-action train env
        self.tok_to_pos = {"s": 0, "rtg": 1, "r": 2}
        self.tok_to_pred_pos = {"s": 2, "rtg": 0, "a": 1, "r": 1}
        self.action_tokenizer = make_tokenizer("minmax", {"vocab_size": config.action_channels,
                                                          "shift": config.discrete_actions})

    @property
    def device(self):
        return self.engine.device

    def forward(self, states=None, actions=None, rewards=None, returns_to_go=None, timesteps=None,
                attention_mask=None, output_hidden_states=None, output_attentions=None, return_dict=None,
                deterministic=True, with_log_probs=False, prompt=None, task_id=None, ddp_kwargs=None,
                context_trjs=None, inference_params=None, past_key_values=None, use_inference_cache=False):
        if not use_inference_cache:
            raise NotImplementedError("only the inference-cache (recurrent) path is implemented")
        if prompt is not None or context_trjs is not None:
            raise NotImplementedError("prompts / retrieval contexts are out of scope")
        cfg = self.config
        B = states.shape[0]
        state_embeds = False
        if states.dim() == 5:
            # image observations [B, T, C, H, W]: /255 + embed_image on the newest frame
            # (online_decision_transformer_model.py:522-526, discrete_decision_transformer_model.py:187-203)
            if self.embed_image is None:
                raise ValueError("image observations need embed_image.* weights in the state_dict")
            states = self.embed_image(states[:, -1].to(self.device)).unsqueeze(1)
            state_embeds = True
        elif states.dim() == 3 and self.img_is_encoded and states.shape[-1] == cfg.d:
            state_embeds = True                     # discrete_decision_transformer_model.py:185-186
        elif states.dim() != 3:
            raise ValueError(f"states must be [B,T,{cfg.state_dim}] or [B,T,C,H,W], got {tuple(states.shape)}")
        elif states.shape[-1] != cfg.state_dim:
            raise ValueError(f"states must be padded to {cfg.state_dim} (DecisionXLSTM.pad_inputs)")
        # is_discrete passed to the head = not actions.is_floating_point(); actions become int64 only for 1-D
        # action envs (online_decision_transformer_model.py:350-352,364)
        discrete = bool(self.is_discrete and actions is not None and actions.shape[-1] == 1)
        s_last = states[:, -1].to(self.device, torch.float32).contiguous()
        g_last = returns_to_go[:, -1].reshape(B).to(self.device, torch.float32).contiguous()
        r_last = None
        if rewards is not None:
            r_last = rewards[:, -1].reshape(B).to(self.device, torch.float32).contiguous()
        if past_key_values is None:
            cache = self.engine.new_state(B)
        elif isinstance(past_key_values, StateCache):
            cache = past_key_values
        else:
            cache = self.engine.new_state(B)
            cache.load_past_key_values(past_key_values)
        flags = (L.XL_FLAG_DISCRETE if discrete else 0) | (L.XL_FLAG_GRAPH if self.use_graph else 0)
        out = self.engine.policy_step(cache, s_last, g_last, r_last, mode=self.mode, flags=flags,
                                      want_logits=True, want_hidden=True, state_embeds=state_embeds)
        tokens = out["action_tokens"].to(torch.long)
        if discrete:
            action_preds = tokens[:, :1].view(B, 1, 1)
            logits = out["action_logits"].view(B, 1, 1, cfg.num_actions)
        else:
            action_preds = out["action_preds"].view(B, 1, cfg.act_dim)
            logits = out["action_logits"].view(B, 1, cfg.act_dim, cfg.num_actions)
        return _Output(last_hidden_state=out["last_hidden_state"], action_preds=action_preds,
                       action_logits=logits, action_tokens=tokens, past_key_values=cache,
                       state_preds=None, return_preds=None, reward_preds=None, action_log_probs=None,
                       hidden_states=None, attentions=None, entropy=None, prompt_infos=None, cross_attentions=None)

    __call__ = forward


@dataclasses.dataclass
class _ReplayBufferDims:
    max_state_dim: Optional[int] = 204
    max_act_dim: Optional[int] = 8


class DiscreteDecisionXLSTM:
    """Agent-side `predict` for ONE env (B == 1), as the reference's rollout loop calls it
    (src/callbacks/evaluation.py:134-139)."""

    def __init__(self, policy: MultiDomainDiscreteDecisionXLSTMModel, use_inference_cache: bool = True,
                 state_mean=None, state_std=None, reset_inf_cache_freq: Optional[int] = None):
        self.policy = policy
        self.device = policy.device
        self.use_inference_cache = use_inference_cache
        self.past_key_values: Optional[Any] = None
        self.replay_buffer = _ReplayBufferDims(policy.config.state_dim, policy.config.act_dim)
        self.state_mean, self.state_std = state_mean, state_std
        self.reset_inf_cache_freq = reset_inf_cache_freq
        self.target_return_type = "predefined"
        self.s_proj_raw = False
        self.ddp_kwargs: Dict[str, Any] = {}

    def pad_inputs(self, states, actions, returns_to_go, timesteps, context_len=5, rewards=None):
        """decision_xlstm.py:11-28 (cache mode): no time padding; zero-pad state -> 204 and action -> 8."""
        if not self.use_inference_cache:
            raise NotImplementedError("only the inference-cache path is implemented")
        attention_mask = torch.ones(actions.shape[1], device=self.device, dtype=torch.long).reshape(1, -1)
        rb = self.replay_buffer
        if rb.max_state_dim is not None and states.dim() == 3 and not self.s_proj_raw:
            pad = rb.max_state_dim - states.shape[-1]
            states = torch.cat([states, torch.zeros((*states.shape[:-1], pad), device=states.device)], dim=-1)
        if rb.max_act_dim is not None and actions.is_floating_point() and states.dim() != 5:
            pad = rb.max_act_dim - actions.shape[-1]
            actions = torch.cat([actions, torch.zeros((*actions.shape[:-1], pad), device=actions.device)], dim=-1)
        return states.float(), actions, returns_to_go.float(), timesteps, attention_mask, rewards

    def predict(self, policy, observation, actions, rewards, returns_to_go, timesteps, state=None,
                episode_start=None, deterministic=True, context_len=5, prompt=None, task_id=None, is_eval=False,
                env_act_dim=None):
        """decision_transformer_sb3.py:621-667."""
        obs_shape, act_dim = observation.shape[1:], actions.shape[-1]
        states = observation.reshape(1, -1, *obs_shape)[:, -context_len:]
        actions = actions.reshape(1, -1, act_dim)[:, -context_len:]
        returns_to_go = returns_to_go.reshape(1, -1, 1)[:, -context_len:]
        timesteps = timesteps.reshape(1, -1)[:, -context_len:]
        if rewards is not None:
            rewards = rewards.reshape(1, -1, 1)[:, -context_len:]
        states, actions, returns_to_go, timesteps, attention_mask, rewards = self.pad_inputs(
            states, actions, returns_to_go, timesteps, context_len=context_len, rewards=rewards)
        if self.state_mean is not None and self.state_std is not None:
            states = (states - self.state_mean) / self.state_std
        a1, a2 = self.get_action_pred(policy, states, actions, rewards, returns_to_go, timesteps, attention_mask,
                                      deterministic, prompt, task_id=task_id, is_eval=is_eval,
                                      env_act_dim=env_act_dim)
        if self.reset_inf_cache_freq is not None:
            cur = int(timesteps[0, -1])
            if cur > 0 and cur % self.reset_inf_cache_freq == 0:
                self.past_key_values = None
        return a1, a2

    def get_action_pred(self, policy, states, actions, rewards, returns_to_go, timesteps, attention_mask,
                        deterministic, prompt, is_eval=False, task_id=None, env_act_dim=None):
        """discrete_decision_transformer_sb3.py:13-72, shared-action-head branch (:60-68)."""
        out = policy(states=states, actions=actions, rewards=rewards, returns_to_go=returns_to_go,
                     timesteps=timesteps, attention_mask=attention_mask, return_dict=True,
                     deterministic=deterministic, prompt=prompt, task_id=task_id, ddp_kwargs=self.ddp_kwargs,
                     use_inference_cache=self.use_inference_cache, past_key_values=self.past_key_values)
        action = out.action_preds[0, -1]
        if self.use_inference_cache:
            self.past_key_values = out.past_key_values
        if env_act_dim is not None:
            action = action[:env_act_dim]
        return action, action
