"""Which TMEM lanes hold a cta_group::1, M = 64 accumulator? Runs the tcgen05 Linear with 64-row tiles under both
readings (xl_set_option gemm_m64_layout 0: rows in lanes 0-15 of every 32-lane quarter; 1: rows in lanes 0-63) against
fp64, then times 128- vs 64-row tiles at the step's proj_up / proj_down shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200.config import preset
from lram_b200.synth import make_state_dict
from lram_b200.engine import XLSTMEngine

cfg = preset("toy")
eng = XLSTMEngine(cfg, make_state_dict(cfg), max_batch=1)
shapes = [(192, 3072, 768), (192, 768, 1536), (64, 2192, 768), (100, 640, 256), (33, 64, 64)]
for layout in (0, 1):
    eng.set_option("gemm_m64_layout", layout)
    eng.set_option("gemm_bm", 64)
    worst = 0.0
    for (M, N, K) in shapes:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g)
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, generator=g)
        ref = (A.double() @ W.double().t() + bias.double()).float()
        out = eng.linear(A.cuda(), W.cuda(), bias.cuda(), None, impl=2)
        torch.cuda.synchronize()
        err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
        worst = max(worst, err)
        print(f"layout={layout} M={M} N={N} K={K} rel_err={err:.3e}", flush=True)
    print(f"layout {layout}: {'CORRECT' if worst < 2e-5 else 'wrong'} (worst {worst:.2e})")
eng.set_option("gemm_m64_layout", 0)
for (M, N, K) in [(192, 3072, 768), (192, 768, 1536), (96, 3072, 768), (384, 5120, 1280)]:
    A = torch.randn(M, K).cuda()
    W = (torch.randn(N, K) * 0.05).to(torch.bfloat16).cuda()
    for bm in (128, 64):
        eng.set_option("gemm_bm", bm)
        for _ in range(5):
            eng.linear(A, W, impl=2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.linear(A, W, impl=2)
        e1.record()
        torch.cuda.synchronize()
        print(f"M={M} N={N} K={K} tile rows {bm}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (split + GEMM, eager)")
eng.close()
