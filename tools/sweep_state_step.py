"""Times the mLSTM state-step kernel alone (xl_mlstm_cell_step) over tilings; prints achieved GB/s on the
algorithmic bytes (C, n, m read+write). GPU box only."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200.config import preset
from lram_b200.synth import make_state_dict
from lram_b200.engine import XLSTMEngine

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="48M")
ap.add_argument("--envs", type=int, default=64)
ap.add_argument("--T", type=int, nargs="+", default=[3, 1])
ap.add_argument("--tilings", default="0x0,1x128,1x64,2x128,4x128,1x32")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--impl", type=int, nargs="+", default=[1, 0])
args = ap.parse_args()
cfg = preset(args.model, num_blocks=1)
B = args.envs
eng = XLSTMEngine(cfg, make_state_dict(cfg), max_batch=B)
NH, DH, inner = cfg.num_heads, cfg.head_dim, cfg.inner
dev = eng.device
# two independent state sets > L2 so that consecutive launches never hit in cache
nset = max(2, int(300e6 // (B * NH * DH * DH * 4)) + 1)
Cs = [torch.randn(B, NH, DH, DH, device=dev) * 0.01 for _ in range(nset)]
ns = [torch.zeros(B, NH, DH, device=dev) for _ in range(nset)]
ms = [torch.zeros(B, NH, device=dev) for _ in range(nset)]
w = torch.zeros(inner, device=dev)
alg = B * (8 * NH * DH * DH + 8 * NH * DH + 8 * NH)
res = []
for impl in args.impl:
    eng.set_option("state_impl", impl)
    for T in args.T:
      qkv = torch.randn(B * T, 3, inner, device=dev) * 0.1
      ig = torch.randn(B * T, NH, device=dev)
      fg = torch.randn(B * T, NH, device=dev) + 3
      for til in args.tilings.split(","):
          rs, cols = (int(x) for x in til.split("x"))
          try:
              for i in range(3):
                  eng.cell_step(Cs[i % nset], ns[i % nset], ms[i % nset], qkv, ig, fg, w, B, T, rs, cols, want_raw=False)
              torch.cuda.synchronize()
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              e0.record()
              for i in range(args.iters):
                  eng.cell_step(Cs[i % nset], ns[i % nset], ms[i % nset], qkv, ig, fg, w, B, T, rs, cols, want_raw=False)
              e1.record()
              torch.cuda.synchronize()
              us = e0.elapsed_time(e1) / args.iters * 1e3
              gbs = alg / (us * 1e-6) / 1e9
              print(f"{args.model} impl={impl} B={B} T={T} rows_split={rs} cols={cols}: {us:8.1f} us  {gbs:7.0f} GB/s  ({gbs / 6539.9:.2f} of measured copy peak)", flush=True)
              res.append(dict(model=args.model, impl=impl, B=B, T=T, rows_split=rs, cols=cols, us=us, gbs=gbs))
          except Exception as e:  # noqa
              print(f"T={T} tiling {til}: {e}")
print(json.dumps(res))
# calibration: what a plain device copy of the same number of bytes achieves on this box
nbytes = B * NH * DH * DH * 4
a = torch.empty(nbytes // 4, device=dev)
bb = [torch.empty_like(a) for _ in range(3)]
for i in range(3):
    bb[i % 3].copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.iters):
    bb[i % 3].copy_(a)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / args.iters * 1e3
print(f"torch copy_ of {nbytes / 1e6:.0f} MB (read+write {2 * nbytes / 1e6:.0f} MB): {us:.1f} us  {2 * nbytes / (us * 1e-6) / 1e9:.0f} GB/s")
