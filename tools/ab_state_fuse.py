"""A/B of the state kernel's finalize variants (xl_set_option "state_fuse"): 0 separate kernel, 1 cluster symmetric,
2 cluster leader, 3 unfused kernel under the cluster shape (measurement aid). Checks that all variants give the same
tokens / hidden states, then times the graph-replayed step."""
import sys, os, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.synth import make_state_dict, make_stream


def run(name, B, modes=(0, 1, 2, 3), steps=100, discrete=False):
    cfg = preset(name)
    sd = make_state_dict(cfg, seed=0)
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    states, rtg, _ = make_stream(cfg, range(B), 4, domains="mixed")
    s_dev = torch.empty(B, cfg.state_dim, device="cuda")
    r_dev = torch.empty(B, device="cuda")
    ref = None
    for f in modes:
        eng.set_option("state_fuse", f)
        cache, out = eng.new_state(B), None
        hid = []
        for t in range(4):
            s_dev.copy_(torch.from_numpy(states[t])); r_dev.copy_(torch.from_numpy(rtg[t]))
            out = eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, want_hidden=True, out=out)
            torch.cuda.synchronize()
            hid.append((out["action_tokens"].clone(), out["last_hidden_state"].clone()))
        if ref is None:
            ref = hid
        else:
            for (t0, h0), (t1, h1) in zip(ref, hid):
                assert torch.equal(t0, t1), (name, f)
                rel = (h0 - h1).abs().max().item() / h0.abs().max().item()
                assert rel < 1e-4, (name, f, rel)
        for _ in range(10):
            eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, want_hidden=True, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.policy_step(cache, s_dev, r_dev, flags=L.XL_FLAG_GRAPH, want_hidden=True, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"model": name, "B": B, "state_fuse": f, "ms_per_step": round(ms, 4),
                          "env_steps_per_s": round(B / ms * 1e3, 1)}), flush=True)
    eng.close()


if __name__ == "__main__":
    run("48M", 64)
    run("110M", 256, steps=30)
    run("206M", 128, steps=30)
    run("16M", 64)
    run("48M", 256, steps=50)
