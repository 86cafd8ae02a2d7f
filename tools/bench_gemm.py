"""Time the hand-written DMMA GEMM (C ABI) on the shapes of the ADC(2)-x matvec."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adcc_b200 import backend as be

PEAK = 37.17  # TFLOP/s, DMMA.8x8x4 issue peak measured by tools/microbench/fp64_peak.cu
shapes = [(784, 16384, 16384, 0), (780, 16384, 16384, 0), (780, 16384, 16384, 128), (8192, 8192, 8192, 0), (16000, 16000, 16000, 0),
          (7800, 8192, 8192, 0), (40, 400, 312000, 0), (1600, 1600, 160000, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) + (0,) for a in sys.argv[1:]]
res = []
for M, N, K, hint in shapes:
    A = torch.randn(M, K, device="cuda", dtype=torch.float64)
    B = torch.randn(N, K, device="cuda", dtype=torch.float64)
    C = be.empty((M, N))
    for _ in range(2):
        be.gemm(A, B, out=C, tile_hint=hint)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); be.gemm(A, B, out=C, tile_hint=hint); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tf = 2.0 * M * N * K / best * 1e-9
    ref = A[:64] @ B[:64].T
    err = (C[:64, :64] - ref).abs().max().item() / ref.abs().max().item()
    print(f"dgemm M{M} N{N} K{K} hint{hint}: {best:.3f} ms {tf:.2f} TFLOP/s frac {tf/PEAK:.3f} relerr {err:.1e}", flush=True)
    res.append(dict(M=M, N=N, K=K, hint=hint, ms=best, tflops=tf))
    del A, B, C
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_gemm.json", "w"))
