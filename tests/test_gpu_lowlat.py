"""GPU parity tests of the small-batch latency path (lram_b200/csrc/xl_lowlat.cu: the whole mLSTM block stack as one
persistent kernel with grid barriers), through the C ABI with `xl_set_option("lowlat", 1)`.

Same bar as tests/test_gpu_parity.py: action tokens bit-exact against the oracle, hidden states / logits within 1e-3
relative. `lowlat_check` fails if a grid barrier of the kernel ever timed out.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from lram_b200 import _lib as L  # noqa: E402
from lram_b200.config import preset  # noqa: E402
from lram_b200.synth import make_state_dict, make_stream  # noqa: E402

REL_TOL = 1e-3


def _engine(name, B, seed=0):
    from lram_b200.engine import XLSTMEngine
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=seed)
    return cfg, sd, XLSTMEngine(cfg, sd, max_batch=B)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


# (model, envs, steps, mode): row buckets 4 / 8 / 16 of the kernel, fused (T = 3) and per-token (T = 1) stepping
@pytest.mark.parametrize("name,B,steps,mode", [("16M", 1, 4, L.XL_MODE_FUSED), ("48M", 2, 3, L.XL_MODE_FUSED),
                                               ("16M", 5, 3, L.XL_MODE_FUSED), ("16M", 3, 2, L.XL_MODE_PER_TOKEN)])
def test_lowlat_policy_step_vs_oracle(name, B, steps, mode):
    from oracle import xlstm_oracle as O
    cfg, sd, eng = _engine(name, B)
    eng.set_option("lowlat", 1)
    ora = O.OraclePolicy(cfg, sd)
    states, rtg, _ = make_stream(cfg, range(B), steps, domains="mixed")
    cache, pkv = eng.new_state(B), None
    eng.launch_count()
    for t in range(steps):
        out = eng.policy_step(cache, torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda(), mode=mode,
                              want_hidden=True, want_logits=True)
        ref = ora.step(torch.from_numpy(states[t]), torch.from_numpy(rtg[t]), past_key_values=pkv)
        pkv = ref["past_key_values"]
        assert torch.equal(out["action_tokens"].cpu().long(), ref["action_tokens"]), f"tokens differ at t={t}"
        assert _rel(out["last_hidden_state"].cpu(), ref["last_hidden_state"]) < REL_TOL, f"hidden at t={t}"
        assert torch.equal(out["action_preds"].cpu(), ref["action_preds"])
        assert _rel(out["action_logits"].cpu().view_as(ref["action_logits"]), ref["action_logits"]) < REL_TOL
    # the persistent kernel really ran: a step is front + ONE stack launch (per token) + head, not 6 per block
    per_step = eng.launch_count() / steps
    assert per_step < 6 * cfg.num_blocks, per_step
    eng.set_option("lowlat_check", 0)
    # recurrent state left behind, in the reference's past_key_values format
    exp = cache.to_past_key_values()
    for i in range(cfg.num_blocks):
        c, n, m = pkv[f"block_{i}"]["mlstm_state"]
        ce, ne, me = exp[f"block_{i}"]["mlstm_state"]
        assert _rel(ce.cpu(), c) < REL_TOL and _rel(ne.cpu(), n) < REL_TOL
        assert (me.cpu() - m).abs().max() < 1e-4
        assert _rel(exp[f"block_{i}"]["conv_state"][0].cpu(), pkv[f"block_{i}"]["conv_state"][0]) < 1e-5
    eng.close()


def test_lowlat_equals_multi_kernel_path_and_graph_replay():
    """No oracle needed: the latency path and the multi-kernel path leave the same state and tokens (GEMV summation
    order is the only difference); CUDA-graph replay of the latency path == its eager launches, bit for bit."""
    B, steps = 2, 5
    cfg, sd, eng = _engine("16M", B)
    states, rtg, _ = make_stream(cfg, range(B), steps, domains="mixed")
    res = {}
    for tag, ll, flags in (("std", 0, 0), ("ll", 1, 0), ("ll_graph", 1, L.XL_FLAG_GRAPH)):
        eng.set_option("lowlat", ll)
        cache = eng.new_state(B)
        s_dev = torch.empty(B, cfg.state_dim, device="cuda")
        r_dev = torch.empty(B, device="cuda")
        out, toks, hids = None, [], []
        for t in range(steps):
            s_dev.copy_(torch.from_numpy(states[t]))
            r_dev.copy_(torch.from_numpy(rtg[t]))
            out = eng.policy_step(cache, s_dev, r_dev, flags=flags, want_hidden=True, out=out)
            torch.cuda.synchronize()
            toks.append(out["action_tokens"].cpu().clone())
            hids.append(out["last_hidden_state"].cpu().clone())
        res[tag] = (torch.stack(toks), torch.stack(hids),
                    [cache.view(i, L.XL_STATE_C).clone() for i in range(cfg.num_blocks)])
    eng.set_option("lowlat_check", 0)
    assert torch.equal(res["ll"][0], res["std"][0])
    assert _rel(res["ll"][1], res["std"][1]) < 1e-4
    for a, b in zip(res["ll"][2], res["std"][2]):
        assert _rel(a, b) < 1e-4
    assert torch.equal(res["ll_graph"][0], res["ll"][0])
    assert torch.equal(res["ll_graph"][1], res["ll"][1])
    for a, b in zip(res["ll_graph"][2], res["ll"][2]):
        assert torch.equal(a, b)
    eng.close()


@pytest.mark.parametrize("name,B,T", [("206M", 1, 3), ("110M", 1, 3), ("206M", 4, 4), ("48M", 8, 1)])
def test_lowlat_encoder_large_shapes_vs_multi_kernel(name, B, T):
    """Largest head sizes / row buckets: encoder step of the latency path vs the multi-kernel path (itself
    oracle-checked in test_gpu_parity.py), two steps so the second one runs on a non-trivial state."""
    cfg, sd, eng = _engine(name, B)
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(B, T, cfg.d, generator=g).cuda() for _ in range(2)]
    outs = {}
    for ll in (0, 1):
        eng.set_option("lowlat", ll)
        cache = eng.new_state(B)
        outs[ll] = [eng.encoder_step(cache, x).cpu() for x in xs]
        outs[ll].append(cache.view(cfg.num_blocks - 1, L.XL_STATE_C).cpu().clone())
    eng.set_option("lowlat_check", 0)
    for a, b in zip(outs[1], outs[0]):
        assert _rel(a, b) < 1e-4
    eng.close()


def test_lowlat_falls_back_when_not_eligible():
    """Shapes outside the kernel's envelope (toy d = 128, or B*T > 16) silently use the multi-kernel path."""
    cfg, sd, eng = _engine("toy128", 3)
    eng.set_option("lowlat", 1)
    cache = eng.new_state(3)
    x = torch.randn(3, 3, cfg.d).cuda()
    eng.launch_count()
    eng.encoder_step(cache, x)
    assert eng.launch_count() >= 4 * cfg.num_blocks      # LN, proj_up (+ pre-cell epilogue), state, finalize, proj_down
    eng.close()


# ------------------------------------------------------------------------------------------------------------
# GEMV-style small-batch kernels (xl_smallm.cu: LN + proj_up + conv/qkv/gate partials in one launch, proj_down in
# another; option "smallm", B*T <= 16 rows): equality with the multi-kernel path they replace (itself oracle-checked
# in test_gpu_parity.py), and proof that they ran.
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,B,mode", [("16M", 1, L.XL_MODE_FUSED), ("48M", 2, L.XL_MODE_FUSED),
                                         ("110M", 5, L.XL_MODE_FUSED), ("16M", 7, L.XL_MODE_PER_TOKEN)])
def test_smallm_front_kernel_equals_three_kernel_path(name, B, mode):
    cfg, sd, eng = _engine(name, B)
    steps = 4
    states, rtg, _ = make_stream(cfg, range(B), steps, domains="mixed")
    res, launches = {}, {}
    for on in (0, 1):
        eng.set_option("smallm", on)
        cache = eng.new_state(B)
        toks, hids = [], []
        eng.launch_count()
        for t in range(steps):
            out = eng.policy_step(cache, torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda(), mode=mode,
                                  want_hidden=True)
            toks.append(out["action_tokens"].cpu().clone())
            hids.append(out["last_hidden_state"].cpu().clone())
        launches[on] = eng.launch_count() / steps
        res[on] = (torch.stack(toks), torch.stack(hids), cache.to_past_key_values())
    assert torch.equal(res[1][0], res[0][0])
    assert _rel(res[1][1], res[0][1]) < 1e-4
    for i in range(cfg.num_blocks):
        for a, b in zip(res[1][2][f"block_{i}"]["mlstm_state"], res[0][2][f"block_{i}"]["mlstm_state"]):
            assert _rel(a.cpu(), b.cpu()) < 1e-4
        assert _rel(res[1][2][f"block_{i}"]["conv_state"][0].cpu(), res[0][2][f"block_{i}"]["conv_state"][0].cpu()) < 1e-5
    # two launches fewer per block and token pass (LN and conv/qkv folded into the GEMV kernel)
    passes = 1 if mode == L.XL_MODE_FUSED else cfg.tokens_per_step
    # (the multi-kernel path itself saves block 0's LayerNorm launch in fused mode: it rides on the embed kernel)
    assert launches[0] - launches[1] >= 2 * cfg.num_blocks * passes - 1, launches
    eng.close()
