"""Odd row pitch (e.g. 5565 = pairs of 106 virtuals): plain-load fallback GEMM kernel vs a re-pitched
copy + the TMA kernel (what backend.gemm does up to REPITCH_MAX elements)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adcc_b200 import backend as be


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for M, N, K in ((16, 5565, 5565), (45, 703, 703), (16, 1225, 1225), (780, 5565, 5565)):
    A = torch.randn(M, K, device="cuda", dtype=torch.float64)
    B = torch.randn(N, K, device="cuda", dtype=torch.float64)
    keep = be.REPITCH_MAX
    be.REPITCH_MAX = 0
    t_fallback = timeit(lambda: be.gemm(A, B))
    ref = be.gemm(A, B)
    be.REPITCH_MAX = keep
    t_repitch = timeit(lambda: be.gemm(A, B))
    err = float((be.gemm(A, B) - ref).abs().max() / ref.abs().max())
    print(f"M={M:4d} N={N:5d} K={K:5d}  fallback {t_fallback:8.3f} ms   re-pitched + TMA {t_repitch:8.3f} ms   "
          f"rel diff {err:.1e}", flush=True)
