"""Standalone check of the tcgen05 Linear against fp64 (run under `timeout` on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200.config import preset
from lram_b200.synth import make_state_dict
from lram_b200.engine import XLSTMEngine

cfg = preset("toy")
eng = XLSTMEngine(cfg, make_state_dict(cfg), max_batch=1)
ok = True
for (M, N, K) in [(128, 64, 64), (128, 128, 128), (192, 3072, 768), (192, 768, 1536), (64, 2192, 768), (3, 512, 256),
                  (300, 274, 128), (768, 4096, 1024)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ref = (A.double() @ W.double().t() + bias.double() + res.double()).float()
    for impl in (1, 2):
        out = eng.linear(A.cuda(), W.cuda(), bias.cuda(), res.cuda(), impl=impl)
        torch.cuda.synchronize()
        err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
        print(f"M={M} N={N} K={K} impl={impl} rel_err={err:.3e}", flush=True)
        ok &= err < 2e-5
    # timing
    for impl in (1, 2):
        a, w = A.cuda(), W.cuda()
        for _ in range(3):
            eng.linear(a, w, impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.linear(a, w, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        print(f"   impl={impl}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (incl. split + host launch)")
print("GEMM CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
