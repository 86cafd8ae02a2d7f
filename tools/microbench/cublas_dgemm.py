"""cuBLAS DGEMM ceiling on this GPU (library reference point only; not used on any product path)."""
import torch, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
for (M, N, K) in [(8192, 8192, 8192), (784, 16384, 16384), (7840, 16384, 8192), (16000, 16000, 16000)]:
    a = torch.randn(M, K, device=dev, dtype=torch.float64)
    b = torch.randn(N, K, device=dev, dtype=torch.float64)
    for _ in range(3):
        c = a @ b.T
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b.T; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tf = 2.0 * M * N * K / best * 1e-9
    print(f"cublas dgemm TN M{M} N{N} K{K}: {best:.3f} ms {tf:.2f} TFLOP/s", flush=True)
    res[f"{M}x{N}x{K}"] = tf
    del a, b, c
# sustained loop 3 s
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
torch.cuda.synchronize()
import time
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
n = 0; t0 = time.time(); e0.record()
while time.time() - t0 < 3.0:
    c = a @ b.T; n += 1
    if n % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
tf = 2.0 * 8192**3 * n / e0.elapsed_time(e1) * 1e-9
print(f"cublas dgemm sustained 8192^3 x{n}: {tf:.2f} TFLOP/s")
res["sustained_8192"] = tf
json.dump(res, open("gpurun_out/cublas_dgemm.json", "w"))
