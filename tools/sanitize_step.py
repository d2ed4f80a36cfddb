"""Small driver for compute-sanitizer (tools/sanitize.sh): runs every kernel family of the hot path once or twice at
small sizes — fused and per-token steps with programmatic dependent launch on, eager and from a CUDA graph, the
one-env GEMV kernels, per-env reset, the chunkwise prefill, an sLSTM stack, the discrete head — and checks the results
against the oracle so that a "clean" sanitizer log belongs to a run that also computed the right thing.

    compute-sanitizer --tool memcheck python tools/sanitize_step.py [case ...]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lram_b200 import _lib as L  # noqa: E402
from lram_b200.config import preset  # noqa: E402
from lram_b200.engine import XLSTMEngine  # noqa: E402
from lram_b200.synth import make_state_dict, make_stream  # noqa: E402
from oracle import xlstm_oracle as O  # noqa: E402  (checker)


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def steps_case(name, B, mode, flags, n=3, discrete=False, opts=None, ring=False):
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=1)
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    for k, v in (opts or {}).items():
        eng.set_option(k, v)
    tok_ring = None
    if ring:
        tok_ring = torch.zeros(4, B, cfg.act_dim, dtype=torch.int32, device="cuda")
        eng.set_token_ring(tok_ring, 0)
    ora = O.OraclePolicy(cfg, sd)
    states, rtg, _ = make_stream(cfg, range(B), n, domains="mixed")
    cache, pkv, out = eng.new_state(B), None, None
    s_dev = torch.empty(B, cfg.state_dim, device="cuda")
    r_dev = torch.empty(B, device="cuda")
    fl = flags | (L.XL_FLAG_DISCRETE if discrete else 0)
    for t in range(n):
        s_dev.copy_(torch.from_numpy(states[t]))
        r_dev.copy_(torch.from_numpy(rtg[t]))
        out = eng.policy_step(cache, s_dev, r_dev, mode=mode, flags=fl, want_hidden=True, want_logits=True, out=out)
        torch.cuda.synchronize()
        ref = ora.step(torch.from_numpy(states[t]), torch.from_numpy(rtg[t]), past_key_values=pkv, discrete=discrete)
        pkv = ref["past_key_values"]
        tok = out["action_tokens"].cpu().long()
        tok = tok[:, :1] if discrete else tok
        assert torch.equal(tok, ref["action_tokens"]), (name, t)
        assert rel(out["last_hidden_state"].cpu(), ref["last_hidden_state"]) < 1e-3
        if tok_ring is not None:
            assert torch.equal(tok_ring[t % 4], out["action_tokens"])
        if t == 1:
            mask = torch.zeros(B, dtype=torch.uint8)
            mask[0] = 1
            eng.reset(cache, mask)
            pkv = O.reset_state_rows(pkv, mask.bool())
    eng.close()


def prefill_case(name, B, Tn, opts=None):
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=2)
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    for k, v in (opts or {}).items():
        eng.set_option(k, v)
    states, rtg, _ = make_stream(cfg, range(B), Tn + 1, domains="mixed")
    st = torch.from_numpy(np.ascontiguousarray(states[:Tn].transpose(1, 0, 2))).cuda()
    rg = torch.from_numpy(np.ascontiguousarray(rtg[:Tn].T)).cuda()
    c1, c2 = eng.new_state(B), eng.new_state(B)
    eng.policy_prefill(c1, st, rg)
    for t in range(Tn):
        eng.policy_step(c2, st[:, t].contiguous(), rg[:, t].contiguous())
    a = eng.policy_step(c1, torch.from_numpy(states[Tn]).cuda(), torch.from_numpy(rtg[Tn]).cuda(), want_hidden=True)
    b = eng.policy_step(c2, torch.from_numpy(states[Tn]).cuda(), torch.from_numpy(rtg[Tn]).cuda(), want_hidden=True)
    torch.cuda.synchronize()
    assert torch.equal(a["action_tokens"], b["action_tokens"])
    assert rel(a["last_hidden_state"], b["last_hidden_state"]) < 1e-3
    eng.close()


def linear_2sm_case():
    """2-SM (cta_group::2) Linear against the 1-SM kernel (bit-identical) and fp64, ragged M and N."""
    cfg = preset("toy")
    eng = XLSTMEngine(cfg, make_state_dict(cfg, seed=1), max_batch=1)
    g = torch.Generator().manual_seed(3)
    for (M, N, K) in [(600, 720, 128), (520, 256, 64)]:
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
        bias, res = torch.randn(N, generator=g).cuda(), torch.randn(M, N, generator=g).cuda()
        eng.set_option("gemm_2cta", 0)
        ref = eng.linear(A, W, bias, res, impl=2)
        for variant in (1, 2):
            eng.set_option("gemm_2cta", variant)
            out = eng.linear(A, W, bias, res, impl=2)
            torch.cuda.synchronize()
            assert torch.equal(out, ref)
        assert rel(out, (A.double() @ W.double().t() + bias.double() + res.double()).float()) < 2e-5
    eng.set_option("gemm_2cta", -1)
    eng.close()


CASES = {
    "fused_eager": lambda: steps_case("toy128", 6, L.XL_MODE_FUSED, 0),
    "fused_graph": lambda: steps_case("toy128", 6, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH),
    "per_token": lambda: steps_case("toy128", 5, L.XL_MODE_PER_TOKEN, 0),
    "one_env": lambda: steps_case("toy128", 1, L.XL_MODE_FUSED, 0),              # GEMV-style small-batch kernels
    "discrete": lambda: steps_case("toy128", 4, L.XL_MODE_FUSED, 0, discrete=True),
    "real_16M": lambda: steps_case("16M", 8, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2),  # DH=256: TMA ring, split-K planes
    "slstm": lambda: steps_case("toy128-ms", 4, L.XL_MODE_FUSED, 0),
    "prefill": lambda: prefill_case("toy128", 2, 32),                            # 96 tokens: tensor-core cell + tail
    # round-2 kernels and options (process-wide GEMM switches are restored by the option's own case ending the process)
    "state_fuse1": lambda: steps_case("16M", 64, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2, opts={"state_fuse": 1}),
    "state_fuse2": lambda: steps_case("16M", 64, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2, opts={"state_fuse": 2}),
    "up_fuse": lambda: steps_case("16M", 50, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2, opts={"up_fuse": 1}),
    "gemm_bm64": lambda: steps_case("16M", 40, L.XL_MODE_FUSED, 0, n=2, opts={"gemm_bm": 64}),
    "gemm_cluster": lambda: steps_case("16M", 40, L.XL_MODE_FUSED, 0, n=2, opts={"gemm_cluster": 2}),
    "token_ring": lambda: steps_case("toy128", 6, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=4, ring=True),
    # tcgen05 chunkwise prefill cell (DH = 256): 132 tokens per env = one full 128-token chunk + a ragged one; batched
    # hi/lo GEMMs, fused chunk update + scan (2 TMEM accumulators, shared-memory store staging), two-level gate scan
    "prefill_tc": lambda: prefill_case("16M", 2, 44),
    "prefill_tc_unfused": lambda: prefill_case("16M", 2, 44, opts={"prefill_tc_fused": 0}),
    "gemm_2sm": linear_2sm_case,
    "small_fuse": lambda: steps_case("16M", 1, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2, opts={"small_state_fuse": 1}),
    # late round 2: packed-fp32 (FFMA2) pre-cell kernel of the step path; the prefill_tc case above now runs the packed
    # persistent sequence conv kernel (named barriers per token run), the single-read prep2 kernel and the side-stream
    # overlap of the S GEMM with the chunk scan
    "conv_pk": lambda: steps_case("16M", 8, L.XL_MODE_FUSED, L.XL_FLAG_GRAPH, n=2, opts={"conv_impl": 2}),
    "conv_pk_toy": lambda: steps_case("toy128", 5, L.XL_MODE_PER_TOKEN, 0, opts={"conv_impl": 2}),
    "prefill_tc_ragged": lambda: prefill_case("16M", 3, 51),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
        print(f"[sanitize_step] {n}: ok", flush=True)
