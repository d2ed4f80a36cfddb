"""GPU tests of the one-env / small-batch kernels (lram_b200/csrc/xl_smallm.cu: LN + proj_up + conv/qkv/gate partials
in one GEMV-style launch, proj_down in another; automatic for B*T <= 4 rows and d <= 1024, option "smallm"), through the
C ABI, plus one-env parity against the oracle on the path the reference's own rollout uses (evaluation.py:80 asserts a
single env).

Same bar as tests/test_gpu_parity.py: action tokens bit-exact against the oracle, hidden states / logits within 1e-3.
(The persistent whole-stack kernel of round 1, xl_lowlat.cu, measured 1.5x slower than this path and was removed in
round 2; its measurements stay in profiles/r01_lowlat_persistent.md.)
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


@pytest.mark.parametrize("name,B,steps,mode", [("16M", 1, 4, L.XL_MODE_FUSED), ("48M", 1, 3, L.XL_MODE_FUSED),
                                               ("110M", 1, 2, L.XL_MODE_FUSED), ("16M", 3, 2, L.XL_MODE_PER_TOKEN)])
def test_small_batch_policy_step_vs_oracle(name, B, steps, mode):
    """One env (the GEMV-style kernels run automatically) and a per-token small batch, graph-replayed, vs the oracle."""
    from oracle import xlstm_oracle as O
    cfg, sd, eng = _engine(name, B)
    ora = O.OraclePolicy(cfg, sd)
    states, rtg, _ = make_stream(cfg, range(B), steps, domains="mixed")
    cache, pkv, out = eng.new_state(B), None, None
    s_dev, r_dev = torch.empty(B, cfg.state_dim, device="cuda"), torch.empty(B, device="cuda")
    for t in range(steps):
        s_dev.copy_(torch.from_numpy(states[t]))
        r_dev.copy_(torch.from_numpy(rtg[t]))
        out = eng.policy_step(cache, s_dev, r_dev, mode=mode, flags=L.XL_FLAG_GRAPH, want_hidden=True, want_logits=True,
                              out=out)
        torch.cuda.synchronize()
        ref = ora.step(torch.from_numpy(states[t]), torch.from_numpy(rtg[t]), past_key_values=pkv)
        pkv = ref["past_key_values"]
        assert torch.equal(out["action_tokens"].cpu().long(), ref["action_tokens"]), f"tokens differ at t={t}"
        assert _rel(out["last_hidden_state"].cpu(), ref["last_hidden_state"]) < REL_TOL, f"hidden at t={t}"
        assert torch.equal(out["action_preds"].cpu(), ref["action_preds"])
        assert _rel(out["action_logits"].cpu().view_as(ref["action_logits"]), ref["action_logits"]) < REL_TOL
    exp = cache.to_past_key_values()
    for i in range(cfg.num_blocks):
        c, n, m = pkv[f"block_{i}"]["mlstm_state"]
        ce, ne, me = exp[f"block_{i}"]["mlstm_state"]
        assert _rel(ce.cpu(), c) < REL_TOL and _rel(ne.cpu(), n) < REL_TOL
        assert (me.cpu() - m).abs().max() < 1e-4
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


@pytest.mark.parametrize("name,B,mode", [("16M", 1, L.XL_MODE_FUSED), ("48M", 1, L.XL_MODE_FUSED), ("48M", 2, L.XL_MODE_FUSED),
                                         ("110M", 5, L.XL_MODE_FUSED), ("206M", 1, L.XL_MODE_FUSED),
                                         ("16M", 7, L.XL_MODE_PER_TOKEN)])
def test_small_batch_cluster_finalize_equals_two_kernel_step(name, B, mode):
    """Few-env path, option "small_state_fuse": the state kernel's cluster (column slabs x row chunks of a head, <= 16
    CTAs) pushes its numerators to rank 0, which finalizes the head -- one launch fewer per block than with the separate
    finalize kernel, same sums in the same order. Tokens are compared with ==; hidden states and the carried state agree
    to 1e-6 (the division by the per-token denominator is folded differently by the compiler in the two kernels)."""
    cfg, sd, eng = _engine(name, B)
    steps = 3
    states, rtg, _ = make_stream(cfg, range(B), steps, domains="mixed")
    res, launches = {}, {}
    for on in (0, 1):
        eng.set_option("smallm", 1)
        eng.set_option("small_state_fuse", on)
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
    assert _rel(res[1][1], res[0][1]) < 1e-6
    for i in range(cfg.num_blocks):
        for a, b in zip(res[1][2][f"block_{i}"]["mlstm_state"], res[0][2][f"block_{i}"]["mlstm_state"]):
            assert _rel(a.cpu(), b.cpu()) < 1e-6 or (a - b).abs().max().item() < 1e-6
    passes = 1 if mode == L.XL_MODE_FUSED else cfg.tokens_per_step
    assert launches[0] - launches[1] == cfg.num_blocks * passes, launches
    eng.close()
