"""tcgen05 Linear with the A tile multicast across a cluster of 2 / 4 column-tile CTAs (xl_set_option gemm_cluster):
correctness against fp64 at the step's shapes (also with split-K planes through the policy step), run under `timeout`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lram_b200.config import preset
from lram_b200.synth import make_state_dict
from lram_b200.engine import XLSTMEngine

cfg = preset("toy")
eng = XLSTMEngine(cfg, make_state_dict(cfg), max_batch=1)
shapes = [(192, 3072, 768), (192, 768, 1536), (64, 2176, 768), (100, 640, 256), (33, 128, 64), (384, 5120, 1280)]
ok = True
for cx in (1, 2, 4):
    eng.set_option("gemm_cluster", cx)
    for (M, N, K) in shapes:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g)
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, generator=g)
        ref = (A.double() @ W.double().t() + bias.double()).float()
        out = eng.linear(A.cuda(), W.cuda(), bias.cuda(), None, impl=2)
        torch.cuda.synchronize()
        err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
        ok &= err < 2e-5
        print(f"cluster={cx} M={M} N={N} K={K} rel_err={err:.3e}", flush=True)
print("GEMM CLUSTER CHECK", "OK" if ok else "FAILED")
eng.close()
sys.exit(0 if ok else 1)
