"""BASELINE.json configs[3]: context prefill of S tokens per env followed by recurrent rollout steps.

    python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 1000

Reports prefill tokens/s (CUDA events, whole call incl. chunk copies), the post-prefill step latency (must equal
the cold-start latency: the recurrent state is O(1) in context), and — with --check N — verifies on a prefix of N
timesteps that prefill leaves the same state as stepping (both through the library). GPU box only.
"""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.synth import make_state_dict, make_stream

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="206M")
ap.add_argument("--envs", type=int, default=1)
ap.add_argument("--tokens", type=int, default=50000)
ap.add_argument("--rollout", type=int, default=1000)
ap.add_argument("--check", type=int, default=64, help="timesteps of the prefix checked against stepping (0 = skip)")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--out", default=None)
ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="xl_set_option passthrough, e.g. prefill_cell=0")
args = ap.parse_args()

cfg = preset(args.model)
sd = make_state_dict(cfg, seed=0)
B = args.envs
Tn = args.tokens // 3
dev = torch.device("cuda", 0)
eng = XLSTMEngine(cfg, sd, max_batch=B, device=dev)
for kv in args.opt:
    name, val = kv.split("=")
    eng.set_option(name, int(val))
n_stream = 256
st_np, rtg_np, _ = make_stream(cfg, range(B), n_stream, domains="mixed")
reps = (Tn + n_stream - 1) // n_stream
states = torch.from_numpy(np.ascontiguousarray(np.tile(st_np, (reps, 1, 1))[:Tn].transpose(1, 0, 2))).to(dev)
rtg = torch.from_numpy(np.ascontiguousarray(np.tile(rtg_np, (reps, 1))[:Tn].T)).to(dev)
res = {"model": args.model, "envs": B, "options": args.opt, "context_tokens": Tn * 3, "context_timesteps": Tn}

if args.check:
    n = min(args.check, Tn)
    a, b = eng.new_state(B), eng.new_state(B)
    eng.policy_prefill(a, states[:, :n].contiguous(), rtg[:, :n].contiguous())
    for t in range(n):
        eng.policy_step(b, states[:, t].contiguous(), rtg[:, t].contiguous())
    torch.cuda.synchronize()
    worst = 0.0
    for i in range(cfg.num_blocks):
        ca, cb = a.view(i, L.XL_STATE_C), b.view(i, L.XL_STATE_C)
        worst = max(worst, ((ca - cb).abs().max() / cb.abs().max().clamp_min(1e-30)).item())
    res["check_timesteps"] = n
    res["check_max_rel_C_diff_vs_stepping"] = worst
    assert worst < 1e-3, worst
    del a, b

cache = eng.new_state(B)
times = []
for r in range(args.reps + 1):
    eng.reset(cache)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.policy_prefill(cache, states, rtg)
    e1.record()
    torch.cuda.synchronize()
    if r > 0:                       # first call allocates the prefill workspace
        times.append(e0.elapsed_time(e1))
ms = statistics.median(times)
res["prefill_ms"] = ms
res["prefill_tokens_per_s"] = B * Tn * 3 / (ms / 1e3)
# algorithmic flops per token: projections 12 d^2 per block (x2 for the bf16 hi/lo passes not counted) + cell
d, Lb, NH, DH = cfg.d, cfg.num_blocks, cfg.num_heads, cfg.head_dim
res["proj_tflops"] = B * Tn * 3 * Lb * 12 * d * d / (ms / 1e3) / 1e12
res["cell_tflops_fp32"] = B * Tn * 3 * Lb * NH * DH * DH * 4 / (ms / 1e3) / 1e12

# rollout after the context: latency must not depend on the context length
s_in = torch.empty(B, cfg.state_dim, device=dev)
r_in = torch.empty(B, device=dev)
out = None
lat = {}
for tag, c in (("after_prefill", cache), ("cold", eng.new_state(B))):
    ev = []
    for t in range(args.rollout + 20):
        s_in.copy_(states[:, t % Tn])
        r_in.copy_(rtg[:, t % Tn])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = eng.policy_step(c, s_in, r_in, flags=L.XL_FLAG_GRAPH, out=out)
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    v = [x.elapsed_time(y) for x, y in ev[20:]]
    lat[tag] = {"p50_ms": statistics.median(v), "p90_ms": sorted(v)[int(0.9 * (len(v) - 1))],
                "env_steps_per_s": B / (statistics.mean(v) / 1e3)}
res["rollout"] = lat
print(json.dumps(res))
if args.out:
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
eng.close()
