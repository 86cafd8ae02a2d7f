"""Round-2 probe (needs a B200): exactness and rate of the CUTLASS tcgen05 INT8 GEMM of this study,
then an FP64 contraction emulated with it (slicing and recombination in torch, unfused) against
adccb_dgemm.  Usage: python tools/studies/int8_gemm/run_probe.py [M N K]"""
import ctypes
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))
sys.path.insert(0, os.path.dirname(HERE))
from build import build                      # noqa: E402
from ozaki_int8_gpu_probe import slice_rows  # noqa: E402


def main():
    M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (784, 8192, 79872)
    lib = ctypes.CDLL(build())
    lib.study_i8_gemm_workspace.restype = ctypes.c_longlong
    lib.study_i8_gemm.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
    dev = torch.device("cuda", 0)
    ws = torch.empty(max(int(lib.study_i8_gemm_workspace(M, N, K)), 16), dtype=torch.uint8, device=dev)

    def i8_gemm(a, b, out=None):
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.int32, device=dev) if out is None else out
        st = lib.study_i8_gemm(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), a.shape[0], b.shape[0],
                               a.shape[1], a.data_ptr(), b.data_ptr(), out.data_ptr(), ws.data_ptr())
        assert st == 0, f"study_i8_gemm failed with status {st}"
        return out

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / reps

    a8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    b8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    c, ms = timed(lambda: i8_gemm(a8, b8))
    ref = torch._int_mm(a8, b8.t().contiguous()) if M > 16 else (a8.int() @ b8.int().t())
    print(f"CUTLASS tcgen05 INT8 GEMM {M}x{N}x{K}: {ms:.3f} ms = {2.0 * M * N * K / ms * 1e-9:.1f} TOP/s, "
          f"exact = {bool(torch.equal(c, ref))}")
    g = torch.Generator(device="cpu").manual_seed(11)
    U = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
    W = ((torch.rand(N, K, generator=g, dtype=torch.float64) - 0.5) * 1e-3).to(dev)
    from adcc_b200 import backend as be
    exact, ms_dmma = timed(lambda: be.gemm(U, W), reps=3)
    print(f"adccb_dgemm (DMMA): {ms_dmma:.2f} ms = {2.0 * M * N * K / ms_dmma * 1e-9:.2f} TFLOP/s")
    # the CUDA slicing kernel against the torch statement of the same arithmetic
    lib.study_slice_rows.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                     ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]
    lib.study_recombine.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                    ctypes.c_void_p, ctypes.c_longlong]
    stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)      # noqa: E731

    def cuda_slices(X, s):
        rows, kk = X.shape
        planes = torch.empty((s, rows, kk), dtype=torch.int8, device=dev)
        expo = torch.empty(rows, dtype=torch.int32, device=dev)
        st = lib.study_slice_rows(stream(), rows, kk, kk, X.data_ptr(), X.stride(0), 7, s, planes.data_ptr(),
                                  expo.data_ptr())
        assert st == 0
        return expo, planes

    for name, X in (("U", U), ("W", W)):
        (expo, planes), ms_sl = timed(lambda: cuda_slices(X, 6), reps=3)
        e_t, planes_t = slice_rows(X, 7, 6)
        same = bool(torch.equal(expo.double().reshape(-1, 1), e_t)) and all(
            torch.equal(planes[p], planes_t[p]) for p in range(6))
        print(f"slice kernel on {name} {tuple(X.shape)}: {ms_sl:.3f} ms, identical to the torch statement = {same}")

    def emulated_cuda(s):
        ea, As = cuda_slices(U, s)
        eb, Bs = cuda_slices(W, s)
        C = torch.empty((M, N), dtype=torch.float64, device=dev)
        tmp = torch.empty((M, N), dtype=torch.int32, device=dev)
        first = True
        for p in range(s):
            for q in range(s - p):
                i8_gemm(As[p], Bs[q], tmp)
                st = lib.study_recombine(stream(), M, N, tmp.data_ptr(), ea.data_ptr(), eb.data_ptr(),
                                         7 * (p + q + 2), 0.0 if first else 1.0, C.data_ptr(), N)
                assert st == 0
                first = False
        return C

    for s in (6, 7):
        C, ms_em = timed(lambda: emulated_cuda(s), reps=2)
        err = float((C - exact).abs().max() / exact.abs().max())
        print(f"emulated FP64 (CUDA slices + tcgen05 INT8 GEMMs + recombine), 7-bit x {s}: {ms_em:.2f} ms "
              f"= {2.0 * M * N * K / ms_em * 1e-9:.1f} TFLOP/s FP64-equivalent, max|C - dmma| / max = {err:.2e}")

    for s in (6, 7):
        t0 = time.perf_counter()
        ea, As = slice_rows(U, 7, s)
        eb, Bs = slice_rows(W, 7, s)
        C = torch.zeros((M, N), dtype=torch.float64, device=dev)
        tmp = torch.empty((M, N), dtype=torch.int32, device=dev)
        for p in range(s):
            for q in range(s - p):
                C += i8_gemm(As[p], Bs[q], tmp).double() * 2.0 ** (-7 * (p + q + 2))
        C = C * torch.exp2(ea) * torch.exp2(eb).t()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        err = float((C - exact).abs().max() / exact.abs().max())
        print(f"emulated FP64, 7-bit x {s} slices: {dt:.1f} ms wall (unfused), slice GEMMs alone "
              f"{s * (s + 1) // 2 * ms:.1f} ms, max|C - dmma| / max = {err:.2e}")


if __name__ == "__main__":
    main()
