"""Debug aid: few-env step with small_state_fuse off / on, element-wise differences of hidden states and state."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.synth import make_state_dict, make_stream

name, B = sys.argv[1].split(":"); B = int(B)
cfg = preset(name); sd = make_state_dict(cfg, seed=0)
states, rtg, _ = make_stream(cfg, range(B), 3, domains="mixed")
res = {}
for on in (0, 1, 1):
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    eng.set_option("smallm", 1); eng.set_option("small_state_fuse", on)
    cache = eng.new_state(B)
    hs = []
    for t in range(1):
        out = eng.policy_step(cache, torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda(), want_hidden=True)
        hs.append(out["last_hidden_state"].cpu().clone())
    pkv = cache.to_past_key_values()
    cur = (hs, pkv)
    if on in res:
        a, b = res[on], cur
        print("on vs on (repeatability): hidden equal", all(torch.equal(x, y) for x, y in zip(a[0], b[0])))
    res[on] = cur
    eng.close()
a, b = res[0], res[1]
for t, (x, y) in enumerate(zip(a[0], b[0])):
    d = (x - y).abs()
    print(f"step {t}: hidden max abs diff {d.max().item():.3e} rel {d.max().item() / x.abs().max().item():.3e} n_diff {(d > 0).sum().item()}/{d.numel()}")
for i in range(cfg.num_blocks):
    for nm, x, y in zip("Cnm", a[1][f"block_{i}"]["mlstm_state"], b[1][f"block_{i}"]["mlstm_state"]):
        d = (x - y).abs()
        if d.max().item() > 0:
            print(f"block {i} {nm}: max diff {d.max().item():.3e} n_diff {(d > 0).sum().item()}/{d.numel()}")
