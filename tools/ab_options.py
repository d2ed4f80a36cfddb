"""A/B of xl_set_option combinations on the graph-replayed policy step: every combination must give the same action
tokens and (to 1e-4) the same hidden states as the first one; then the step is timed with CUDA events.

    python tools/ab_options.py 48M:64 "state_fuse=0" "state_fuse=2" "up_fuse=0" ...
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lram_b200 import _lib as L  # noqa: E402
from lram_b200.config import preset  # noqa: E402
from lram_b200.engine import XLSTMEngine  # noqa: E402
from lram_b200.synth import make_state_dict, make_stream  # noqa: E402


def run(name, B, combos, steps=100, nsteps_check=4):
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=0)
    states, rtg, _ = make_stream(cfg, range(B), nsteps_check, domains="mixed")
    s_dev = torch.empty(B, cfg.state_dim, device="cuda")
    r_dev = torch.empty(B, device="cuda")
    ref = None
    for combo in combos:
        eng = XLSTMEngine(cfg, sd, max_batch=B)
        for kv in combo.split(","):
            if kv:
                k, v = kv.split("=")
                eng.set_option(k, int(v))
        cache, out = eng.new_state(B), None
        hid = []
        for t in range(nsteps_check):
            s_dev.copy_(torch.from_numpy(states[t]))
            r_dev.copy_(torch.from_numpy(rtg[t]))
            out = eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, want_hidden=True, out=out)
            torch.cuda.synchronize()
            hid.append((out["action_tokens"].clone(), out["last_hidden_state"].clone()))
        worst = 0.0
        if ref is None:
            ref = hid
        else:
            for (t0, h0), (t1, h1) in zip(ref, hid):
                assert torch.equal(t0, t1), (name, combo)
                worst = max(worst, (h0 - h1).abs().max().item() / h0.abs().max().item())
            assert worst < 1e-4, (name, combo, worst)
        for _ in range(10):
            eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for _rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        print(json.dumps({"model": name, "B": B, "options": combo, "ms_per_step": round(best, 4),
                          "env_steps_per_s": round(B / best * 1e3, 1), "launches_per_step": eng.launch_count() // (3 * steps),
                          "max_rel_hidden_vs_first": float(f"{worst:.2e}")}), flush=True)
        eng.close()


if __name__ == "__main__":
    model, B = sys.argv[1].split(":")
    run(model, int(B), sys.argv[2:] or [""], steps=int(os.environ.get("AB_STEPS", "100")))
