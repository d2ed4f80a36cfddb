"""CPU tests of the checkpoint reader (SB3 zip layout, key fix-ups, config inference) and of the image front-end
(product ImpalaCNN module == the oracle's functional restatement). GPU tests at the bottom run the image / embedded
paths and a checkpoint round trip through the C-ABI."""
import os

import numpy as np
import pytest
import torch

from lram_b200 import _lib as L
from lram_b200.checkpoint import fix_policy_keys, infer_config, load_policy_zip, save_policy_zip
from lram_b200.config import preset
from lram_b200.image_encoder import ImpalaCNN, make_impala_state_dict, split_image_weights
from lram_b200.synth import make_state_dict, make_stream


@pytest.mark.parametrize("prefix", ["", "module.", "_orig_mod.", "module._orig_mod."])
def test_sb3_zip_roundtrip_and_key_fixups(tmp_path, prefix):
    cfg = preset("toy")
    sd = make_state_dict(cfg, seed=3)
    sd["predict_state.weight"] = torch.zeros(4, cfg.d)      # dropped by default (load_state_head=False)
    sd["predict_state.bias"] = torch.zeros(4)
    mean, std = torch.arange(204.0), torch.ones(204) * 2
    path = str(tmp_path / "model.zip")
    save_policy_zip(path, sd, state_mean=mean, state_std=std, data={"num_timesteps": 7}, prefix=prefix)
    got, variables = load_policy_zip(path)
    assert "predict_state.weight" not in got and "predict_state.bias" not in got
    want = {k: v for k, v in sd.items() if not k.startswith("predict_state")}
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(variables["state_mean"], mean) and torch.equal(variables["state_std"], std)
    # without the action head (load_kwargs.load_action_head=False)
    got2, _ = load_policy_zip(path, load_action_head=False)
    assert "action_net.0.weight" not in got2 and "embed_state.weight" in got2


def test_bare_state_dict_file_and_legacy_head_names(tmp_path):
    cfg = preset("toy")
    sd = make_state_dict(cfg, seed=1)
    path = str(tmp_path / "policy.pt")
    torch.save(sd, path)
    got, variables = load_policy_zip(path)
    assert variables == {} and torch.equal(got["embed_ln.weight"], sd["embed_ln.weight"])
    legacy = fix_policy_keys({"mu.weight": torch.ones(2, 2), "mu.bias": torch.ones(2),
                              "log_std.weight": torch.ones(2, 2), "log_std.bias": torch.ones(2)})
    assert set(legacy) == {"mu.0.weight", "mu.0.bias", "log_std.0.weight", "log_std.0.bias"}


@pytest.mark.parametrize("name", ["toy", "toy128", "16M", "48M", "110M", "206M"])
def test_infer_config_from_shapes(name):
    cfg = preset(name)
    if name in ("toy", "toy128"):
        sd = make_state_dict(cfg, seed=0)
    else:  # shapes only: meta tensors, no memory
        with torch.device("meta"):
            sd = make_state_dict_meta(cfg)
    got = infer_config(sd)
    for f in ("embedding_dim", "num_blocks", "num_heads", "conv1d_kernel_size", "qkv_proj_blocksize", "state_dim",
              "act_dim", "inner", "head_dim", "head_out"):
        assert getattr(got, f) == getattr(cfg, f), f


def make_state_dict_meta(cfg):
    d, inner, nh, ks, bs = cfg.d, cfg.inner, cfg.num_heads, cfg.conv1d_kernel_size, cfg.qkv_proj_blocksize
    e = lambda *s: torch.empty(*s)
    sd = {"embed_state.weight": e(d, cfg.state_dim), "action_net.0.weight": e(cfg.head_out, d)}
    for i in range(cfg.num_blocks):
        p = f"encoder.layers.blocks.{i}.xlstm."
        sd[p + "proj_up.weight"] = e(2 * inner, d)
        sd[p + "mlstm_cell.igate.weight"] = e(nh, 3 * inner)
        sd[p + "conv1d.conv.weight"] = e(inner, 1, ks)
        sd[p + "q_proj.weight"] = e(inner // bs, bs, bs)
    return sd


@pytest.mark.parametrize("name", ["toy-ms", "toy128-ms", "48M-ms"])
def test_infer_config_recognises_slstm_blocks(name):
    """xLSTM[a:b] checkpoints: sLSTM positions and the feed-forward width come back from the key set / shapes."""
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=0) if name != "48M-ms" else _meta_state_dict(cfg)
    got = infer_config(sd)
    assert got.slstm_at == cfg.slstm_at and got.ffn_dim == cfg.ffn_dim
    assert (got.d, got.num_blocks, got.num_heads, got.inner) == (cfg.d, cfg.num_blocks, cfg.num_heads, cfg.inner)


def _meta_state_dict(cfg):
    """shapes only (no 48M of random numbers on the CPU test path)"""
    sd = make_state_dict(preset("toy-ms"), seed=0)
    out = {}
    d, inner, nh, ks, bs, ff = cfg.d, cfg.inner, cfg.num_heads, cfg.conv1d_kernel_size, cfg.qkv_proj_blocksize, cfg.ffn_dim
    e = lambda *s: torch.empty(*s)
    for k, v in sd.items():
        if not k.startswith("encoder.layers.blocks."):
            out[k] = v
    out["embed_state.weight"] = e(d, cfg.state_dim)
    out["action_net.0.weight"] = e(cfg.head_out, d)
    for i in range(cfg.num_blocks):
        p = f"encoder.layers.blocks.{i}."
        if cfg.is_slstm(i):
            out[p + "xlstm.slstm_cell._recurrent_kernel_"] = e(nh, d // nh, 4, d // nh)
            out[p + "ffn.proj_down.weight"] = e(d, ff)
            out[p + "ffn.proj_up.weight"] = e(2 * ff, d)
            continue
        out[p + "xlstm.proj_up.weight"] = e(2 * inner, d)
        out[p + "xlstm.mlstm_cell.igate.weight"] = e(nh, 3 * inner)
        out[p + "xlstm.conv1d.conv.weight"] = e(inner, 1, ks)
        out[p + "xlstm.q_proj.weight"] = e(inner // bs, bs, bs)
    return out


def test_infer_config_rejects_uncovered_variants():
    cfg = preset("toy")
    sd = make_state_dict(cfg, seed=0)
    bad = dict(sd)
    bad["encoder.layers.blocks.1.ffn.proj_up.weight"] = torch.zeros(1)
    with pytest.raises(NotImplementedError, match="feed-forward"):
        infer_config(bad)
    bad = dict(sd)
    bad["encoder.layers.blocks.0.xlstm_norm.bias"] = torch.zeros(cfg.d)
    with pytest.raises(NotImplementedError, match="ln_bias"):
        infer_config(bad)
    with pytest.raises(ValueError):
        infer_config({"embed_state.weight": torch.zeros(4, 4)})


def test_impala_module_matches_oracle_restatement():
    """product nn.Module (reference key names) == the oracle's functional form, on seeded frames."""
    from oracle.xlstm_oracle import OraclePolicy
    cfg = preset("toy")
    sd = make_state_dict(cfg, seed=0)
    sd.update(make_impala_state_dict(cfg.d, (3, 64, 64), seed=5))
    assert sd["embed_image.linear.0.weight"].shape == (cfg.d, 2048)          # 32 x 8 x 8 after three /2 pools
    assert sd["embed_image.cnn.0.conv.weight"].shape == (16, 3, 3, 3)
    assert sd["embed_image.cnn.2.residual_1.conv_1.weight"].shape == (32, 32, 3, 3)
    net = ImpalaCNN((3, 64, 64), cfg.d).eval()
    net.load_state_dict(split_image_weights(sd))
    frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (3, 3, 64, 64), dtype=np.uint8))
    from lram_b200.image_encoder import scale_frames
    a = net(scale_frames(frames))
    b = OraclePolicy(cfg, sd).embed_image(frames)
    assert a.shape == (3, cfg.d) and torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert torch.equal(net(scale_frames(frames.float())), a)                 # raw 0..255 floats scale the same way
    with pytest.raises(TypeError):
        net(frames)                                                          # the module never guesses the scaling
    assert (a >= 0).all() and a.abs().max() > 0                              # out_relu


# ------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
def test_checkpoint_to_engine_roundtrip_gpu(tmp_path):
    """zip -> load_policy_zip -> infer_config -> engine == engine built from the in-memory state_dict."""
    from lram_b200.engine import XLSTMEngine
    cfg = preset("toy128")
    sd = make_state_dict(cfg, seed=2)
    path = str(tmp_path / "agent.zip")
    save_policy_zip(path, sd, prefix="module._orig_mod.")
    sd2, _ = load_policy_zip(path)
    cfg2 = infer_config(sd2)
    B = 4
    e1, e2 = XLSTMEngine(cfg, sd, max_batch=B), XLSTMEngine(cfg2, sd2, max_batch=B)
    states, rtg, _ = make_stream(cfg, range(B), 3, domains="mixed")
    c1, c2 = e1.new_state(B), e2.new_state(B)
    for t in range(3):
        s, g = torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda()
        o1 = e1.policy_step(c1, s, g, want_hidden=True)
        o2 = e2.policy_step(c2, s, g, want_hidden=True)
        assert torch.equal(o1["action_tokens"], o2["action_tokens"])
        assert torch.equal(o1["last_hidden_state"], o2["last_hidden_state"])
    e1.close(); e2.close()


@gpu
@pytest.mark.parametrize("discrete", [True, False])
def test_image_observations_vs_oracle_gpu(discrete):
    """Atari/Procgen-style rollout step: uint8 frames -> ImpalaCNN (PyTorch, GPU) -> state embeddings -> CUDA path
    vs the oracle's CPU restatement of the same pipeline. Tokens bit-exact, hidden within 1e-3."""
    from lram_b200.decision_xlstm import MultiDomainDiscreteDecisionXLSTMModel
    from oracle.xlstm_oracle import OraclePolicy
    cfg = preset("toy128")
    sd = make_state_dict(cfg, seed=0)
    sd.update(make_impala_state_dict(cfg.d, (3, 64, 64), seed=9))
    B, steps = 5, 3
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model = MultiDomainDiscreteDecisionXLSTMModel(cfg, sd, max_batch=B)
        ora = OraclePolicy(cfg, sd)
        rng = np.random.default_rng(3)
        pkv_gpu, pkv = None, None
        for t in range(steps):
            frames = torch.from_numpy(rng.integers(0, 256, (B, 1, 3, 64, 64), dtype=np.uint8))
            rtg = torch.full((B, 1, 1), 5.0 - 0.1 * t)
            act = torch.zeros(B, 1, 1, dtype=torch.long) if discrete else torch.zeros(B, 1, cfg.act_dim)
            out = model(states=frames, actions=act, rewards=None, returns_to_go=rtg,
                        timesteps=torch.zeros(B, 1, dtype=torch.long), past_key_values=pkv_gpu,
                        use_inference_cache=True)
            pkv_gpu = out.past_key_values
            ref = ora.step(frames[:, 0], rtg.view(B), past_key_values=pkv, discrete=discrete)
            pkv = ref["past_key_values"]
            tok = out.action_tokens.cpu()
            if discrete:
                assert torch.equal(tok[:, :1], ref["action_tokens"])
                assert torch.equal(out.action_preds.cpu().view(B, 1), ref["action_tokens"])
            else:
                assert torch.equal(tok, ref["action_tokens"])
                assert torch.equal(out.action_preds.cpu().view(B, -1), ref["action_preds"])
            hid, rh = out.last_hidden_state.cpu(), ref["last_hidden_state"]
            assert (hid - rh).abs().max().item() <= 1e-3 * rh.abs().max().item()
        model.engine.close()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@gpu
def test_state_embeds_flag_and_host_rejection_gpu():
    """XL_FLAG_STATE_EMBEDS with embeddings == the oracle fed the same embeddings (the reference's img_is_encoded
    inputs); the host-buffer entry point refuses the flag."""
    from lram_b200.engine import XLSTMEngine
    from oracle.xlstm_oracle import OraclePolicy
    cfg = preset("toy128")
    sd = make_state_dict(cfg, seed=4)
    B = 6
    eng, ora = XLSTMEngine(cfg, sd, max_batch=B), OraclePolicy(cfg, sd)
    g = torch.Generator().manual_seed(0)
    cache, pkv = eng.new_state(B), None
    for t in range(3):
        emb = torch.randn(B, cfg.d, generator=g) * 0.5
        rtg = torch.rand(B, generator=g) * 10
        out = eng.policy_step(cache, emb.cuda(), rtg.cuda(), want_hidden=True, state_embeds=True)
        ref = ora.step(emb, rtg, past_key_values=pkv, state_embeds=True)
        pkv = ref["past_key_values"]
        assert torch.equal(out["action_tokens"].cpu().long(), ref["action_tokens"])
        rh = ref["last_hidden_state"]
        assert (out["last_hidden_state"].cpu() - rh).abs().max().item() <= 1e-3 * rh.abs().max().item()
    h_s = torch.zeros(B, cfg.state_dim).pin_memory()
    h_g = torch.zeros(B).pin_memory()
    h_t = torch.zeros(B, cfg.act_dim, dtype=torch.int32).pin_memory()
    h_a = torch.zeros(B, cfg.act_dim).pin_memory()
    with pytest.raises(RuntimeError, match="state embeddings"):
        eng.policy_step_host(cache, h_s, h_g, h_t, h_a, flags=L.XL_FLAG_STATE_EMBEDS)
    eng.close()


@gpu
@pytest.mark.parametrize("name,B", [("48M", 16), ("16M", 64), ("toy128", 9)])
def test_splitk_planes_equal_unsplit_within_rounding_gpu(name, B):
    """Split-K projections (planes added by the consumer kernels) vs the unsplit GEMMs: same tokens, hidden states
    equal to fp32 rounding (the summation order over K differs), and run-to-run bit reproducible."""
    from lram_b200.engine import XLSTMEngine
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=0)
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    states, rtg, _ = make_stream(cfg, range(B), 3, domains="mixed")

    def run(splitk):
        eng.set_option("gemm_splitk", splitk)
        cache = eng.new_state(B)
        outs = []
        for t in range(3):
            o = eng.policy_step(cache, torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda(),
                                want_hidden=True)
            outs.append((o["action_tokens"].clone(), o["last_hidden_state"].clone()))
        return outs

    a, b, a2 = run(8), run(1), run(8)
    for (ta, ha), (tb, hb), (tc, hc) in zip(a, b, a2):
        assert torch.equal(ta, tb) and torch.equal(ta, tc)
        assert torch.equal(ha, hc)                                           # deterministic
        assert (ha - hb).abs().max().item() <= 2e-5 * hb.abs().max().item()
    eng.close()
