"""Step latency at small env batches: three-kernel front (LN, tcgen05 proj_up, conv/qkv) vs the fused GEMV-style front
kernel (xl_smallm.cu). Graph-replayed xl_policy_step, device-resident inputs, CUDA events per step. GPU box only.

    python tools/bench_smallm.py [--models 16M,48M,110M,206M] [--envs 1,2,4,5] [--steps 300] [--out FILE]
"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.synth import make_state_dict, make_stream

ap = argparse.ArgumentParser()
ap.add_argument("--models", default="16M,48M,110M,206M")
ap.add_argument("--envs", default="1,4")
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = torch.device("cuda", 0)
rows = []
for name in args.models.split(","):
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=0)
    for B in [int(b) for b in args.envs.split(",")]:
        eng = XLSTMEngine(cfg, sd, max_batch=B, device=dev)
        st, rtg, _ = make_stream(cfg, range(B), 64, domains="mixed")
        st, rtg = torch.from_numpy(st).to(dev), torch.from_numpy(rtg).to(dev)
        s_in, r_in = torch.empty(B, cfg.state_dim, device=dev), torch.empty(B, device=dev)
        toks = {}
        for on in (0, 1):
            eng.set_option("smallm", on)
            cache, out, ev, tk = eng.new_state(B), None, [], []
            for t in range(args.steps + 20):
                s_in.copy_(st[t % 64]); r_in.copy_(rtg[t % 64])
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                out = eng.policy_step(cache, s_in, r_in, flags=L.XL_FLAG_GRAPH, out=out)
                b.record()
                ev.append((a, b))
                if t < 40:
                    tk.append(out["action_tokens"].clone())
            torch.cuda.synchronize()
            v = [x.elapsed_time(y) for x, y in ev[20:]]
            toks[on] = torch.stack(tk).cpu()
            p50 = statistics.median(v)
            by = B * cfg.state_bytes_per_env() * 2 + 2 * (cfg.encoder_params() + cfg.d * 256 + cfg.head_out * cfg.d)
            rows.append({"model": name, "envs": B, "smallm": on, "p50_us": p50 * 1e3,
                         "p90_us": sorted(v)[int(0.9 * (len(v) - 1))] * 1e3, "env_steps_per_s": B / (statistics.mean(v) / 1e3),
                         "compulsory_MB": by / 1e6, "achieved_GBps": by / (p50 * 1e-3) / 1e9})
            print(json.dumps(rows[-1]), flush=True)
        rows.append({"model": name, "envs": B, "tokens_equal_first_40_steps": bool(torch.equal(toks[0], toks[1]))})
        print(json.dumps(rows[-1]), flush=True)
        eng.close()
if args.out:
    with open(args.out, "w") as fh:
        json.dump(rows, fh, indent=1)
