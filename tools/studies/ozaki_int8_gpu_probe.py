"""Feasibility probe for round 2 (study only, NOT the product path and not a benchmark arm): an
Ozaki-type FP64 GEMM emulated with INT8 slice products through ``torch._int_mm`` (cuBLASLt IMMA),
to measure on a B200 (a) the INT8 GEMM rate at the headline shapes and (b) the end-to-end error and
time of the emulated contraction against the library's own DMMA GEMM.  If the numbers hold, the
slice products become a hand-written tcgen05 ``kind::i8`` kernel behind the C ABI.

Runs on CPU too (tiny shapes) so that its logic is testable without a GPU:
    python tools/studies/ozaki_int8_gpu_probe.py 64 96 4096        # M N K
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def slice_rows(X, beta, s):
    """Row exponents e and s int8 planes with X ~= 2^e * sum_p 2^(-beta (p+1)) X_p (|X_p| < 2^beta)."""
    amax = X.abs().amax(dim=1, keepdim=True).clamp_min(torch.finfo(X.dtype).tiny)
    e = torch.frexp(amax)[1].to(X.dtype)            # amax = f * 2^e, 0.5 <= f < 1: |X / 2^e| < 1 strictly
    R = X * torch.exp2(-e)
    planes = []
    for _ in range(s):
        R = R * float(2 ** beta)
        P = torch.trunc(R)
        planes.append(P.to(torch.int8))
        R = R - P
    return e, planes


def emulated_gemm(A, B, beta=7, s=6, k_chunk=1 << 16):
    """C = A B^T from s (s + 1) / 2 exact INT8 slice products (INT32 accumulation, chunked over K so
    that 2^(2 beta) * k_chunk < 2^31), combined in FP64."""
    ea, As = slice_rows(A, beta, s)
    eb, Bs = slice_rows(B, beta, s)
    M, K = A.shape
    C = torch.zeros((M, B.shape[0]), dtype=torch.float64, device=A.device)
    for p in range(s):
        for q in range(s - p):
            acc = torch.zeros_like(C)
            for k0 in range(0, K, k_chunk):
                k1 = min(K, k0 + k_chunk)
                acc += torch._int_mm(As[p][:, k0:k1].contiguous(), Bs[q][:, k0:k1].t().contiguous()).double()
            C += acc * 2.0 ** (-beta * (p + q + 2))
    return C * torch.exp2(ea) * torch.exp2(eb).t()


def main():
    M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (784, 8192, 79800)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator(device="cpu").manual_seed(11)
    # _int_mm wants M > 16 and K, N multiples of 8 on CUDA: pad K with zeros
    Kp = (K + 7) // 8 * 8
    U = torch.zeros(M, Kp, dtype=torch.float64)
    W = torch.zeros(N, Kp, dtype=torch.float64)
    U[:, :K] = torch.randn(M, K, generator=g, dtype=torch.float64)
    W[:, :K] = (torch.rand(N, K, generator=g, dtype=torch.float64) - 0.5) * 1e-3
    U, W = U.to(dev), W.to(dev)

    def timed(fn, reps=3):
        fn()
        if dev == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        if dev == "cuda":
            torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) / reps

    ref, t_ref = timed(lambda: U @ W.t())
    if dev == "cuda":
        from adcc_b200 import backend as be
        own, t_own = timed(lambda: be.gemm(U, W))
        print(f"adccb_dgemm (DMMA): {t_own * 1e3:8.2f} ms = {2.0 * M * N * K / t_own * 1e-12:6.2f} TFLOP/s, "
              f"max|own - torch| / max = {float((own - ref).abs().max() / ref.abs().max()):.1e}")
    a8 = torch.randint(-127, 128, (M, Kp), dtype=torch.int8, device=dev)
    b8 = torch.randint(-127, 128, (Kp, N), dtype=torch.int8, device=dev)
    _, t_i8 = timed(lambda: torch._int_mm(a8, b8))
    print(f"library FP64 GEMM : {t_ref * 1e3:8.2f} ms = {2.0 * M * N * K / t_ref * 1e-12:6.2f} TFLOP/s")
    print(f"one INT8 slice GEMM: {t_i8 * 1e3:8.2f} ms = {2.0 * M * N * Kp / t_i8 * 1e-12:6.1f} TOP/s")
    exact = ref
    if M * N <= 1 << 14:                                  # small enough for an exact long-double check
        import numpy as np
        exact = torch.from_numpy((U.cpu().numpy().astype(np.longdouble)
                                  @ W.cpu().numpy().astype(np.longdouble).T).astype(np.float64)).to(dev)
    for s in (5, 6, 7):
        C, t = timed(lambda: emulated_gemm(U, W, 7, s), reps=1)
        err = float((C - exact).abs().max() / exact.abs().max())
        print(f"emulated, 7-bit x {s} slices ({s * (s + 1) // 2:2d} slice GEMMs, unfused torch ops): "
              f"{t * 1e3:9.2f} ms, max error / max|C| = {err:.2e}  "
              f"(slice GEMMs alone would take {s * (s + 1) // 2 * t_i8 * 1e3:.1f} ms)")


if __name__ == "__main__":
    main()
