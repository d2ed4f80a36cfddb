"""Times the mLSTM state-stream kernel ALONE (CUDA events around each of its launches inside
xl_mlstm_cell_step, via xl_profile_begin/end) over implementations and tilings; prints achieved GB/s on the
algorithmic bytes (C, n, m read+write). GPU box only.

    python tools/sweep_state_step.py --model 48M --envs 64 --variants "2:0:0:0,2:4:6:1,1:0:0:0"

A variant is impl:rows_split:stages:ctas_per_sm (0 = automatic / default).
"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200.config import preset
from lram_b200.synth import make_state_dict
from lram_b200.engine import XLSTMEngine

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="48M")
ap.add_argument("--envs", type=int, default=64)
ap.add_argument("--T", type=int, nargs="+", default=[3])
ap.add_argument("--variants", default="2:0:0:0,1:0:0:0,0:0:0:0")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--peak", type=float, default=6650.0)
ap.add_argument("--out", default=None)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--d", type=int, default=0, help="override embedding_dim (DH = d/2)")
args = ap.parse_args()
cfg = preset(args.model, num_blocks=1, **({"embedding_dim": args.d} if args.d else {}))
B = args.envs
eng = XLSTMEngine(cfg, make_state_dict(cfg), max_batch=B)
NH, DH, inner = cfg.num_heads, cfg.head_dim, cfg.inner
dev = eng.device
for kv in args.opt:
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
# several independent state sets > L2 so that consecutive launches never hit in cache
nset = max(2, int(300e6 // (B * NH * DH * DH * 4)) + 1)
Cs = [torch.randn(B, NH, DH, DH, device=dev) * 0.01 for _ in range(nset)]   # layout is irrelevant for timing
ns = [torch.zeros(B, NH, DH, device=dev) for _ in range(nset)]
ms = [torch.zeros(B, NH, device=dev) for _ in range(nset)]
w = torch.zeros(inner, device=dev)
alg = B * (8 * NH * DH * DH + 8 * NH * DH + 8 * NH)
res = []


def timed(T, rs, qkv, ig, fg):
    for i in range(3):
        eng.cell_step(Cs[i % nset], ns[i % nset], ms[i % nset], qkv, ig, fg, w, B, T, rs, 0, want_raw=False, slab=True)
    torch.cuda.synchronize()
    eng.lib.xl_profile_begin(eng.handle)
    for i in range(args.iters):
        eng.cell_step(Cs[i % nset], ns[i % nset], ms[i % nset], qkv, ig, fg, w, B, T, rs, 0, want_raw=False, slab=True)
    torch.cuda.synchronize()
    ms_sum, cnt, step_ms = C.c_double(), C.c_int64(), C.c_double()
    eng.lib.xl_profile_end(eng.handle, C.byref(ms_sum), C.byref(cnt), C.byref(step_ms))
    return ms_sum.value / cnt.value * 1e3


for T in args.T:
    qkv = torch.randn(B * T, 3, inner, device=dev) * 0.1
    ig = torch.randn(B * T, NH, device=dev)
    fg = torch.randn(B * T, NH, device=dev) + 3
    for var in args.variants.split(","):
        impl, rs, stages, cps = (int(x) for x in var.split(":"))
        try:
            eng.set_option("state_impl", impl)
            eng.set_option("state_stages", stages)
            eng.set_option("state_ctas_per_sm", cps)
            us = timed(T, rs, qkv, ig, fg)
            gbs = alg / (us * 1e-6) / 1e9
            print(f"{args.model} B={B} T={T} impl={impl} rows_split={rs} stages={stages} ctas/SM={cps}: "
                  f"{us:8.1f} us  {gbs:7.0f} GB/s  ({gbs / args.peak:.3f} of {args.peak:.0f})", flush=True)
            res.append(dict(model=args.model, B=B, T=T, impl=impl, rows_split=rs, stages=stages, ctas_per_sm=cps,
                            us=us, gbs=gbs))
        except Exception as e:  # noqa
            print(f"T={T} variant {var}: {e}")
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
print(f"torch copy_ of {nbytes / 1e6:.0f} MB (read+write {2 * nbytes / 1e6:.0f} MB): {us:.1f} us  "
      f"{2 * nbytes / (us * 1e-6) / 1e9:.0f} GB/s")
res.append(dict(model=args.model, B=B, torch_copy_us=us, torch_copy_gbs=2 * nbytes / (us * 1e-6) / 1e9))
# the same copy timed the way the state kernel is timed above: one event pair per launch, other work between
evs = []
for i in range(args.iters):
    ns[0].add_(1.0)                                    # an unrelated small kernel between launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bb[i % 3].copy_(a)
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
us2 = sum(x.elapsed_time(y) for x, y in evs) / args.iters * 1e3
print(f"  same copy, one event pair per launch: {us2:.1f} us  {2 * nbytes / (us2 * 1e-6) / 1e9:.0f} GB/s")
res.append(dict(model=args.model, B=B, torch_copy_per_launch_us=us2))
if args.out:
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
