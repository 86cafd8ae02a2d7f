"""Numerical study (CPU, NumPy; nothing here is on the product path): how many INT8 slices would an
Ozaki-type FP64 GEMM emulation need for the headline contraction to meet the 1e-11 sigma tolerance?

Model of the scheme (Ozaki / Ootomo / Yokota): every row of A and of B is scaled by a power of two,
truncated into `s` signed slices of `beta` bits, the slice products A_p B_q^T are EXACT in INT32 for
K <= 2^31 / 2^(2 beta), and C ~= sum_{p+q < s} 2^(-beta (p+q+2)) A_p B_q^T (s (s+1) / 2 slice GEMMs).
Operands here have the statistics of the vvvv term: U ~ N(0,1) trial vector rows, W uniform in
+-scale (the counter-based hash ERIs), K = 79800 canonical virtual pairs.
"""
import sys

import numpy as np


def slices(X, beta, s):
    """X[r, :] = 2^e[r] * sum_p 2^(-beta (p+1)) X_p[r, :] + remainder, X_p integer in (-2^beta, 2^beta)."""
    e = np.frexp(np.max(np.abs(X), axis=1, keepdims=True))[1].astype(float)   # max = f 2^e, 0.5 <= f < 1
    R = X / 2.0 ** e                                   # |R| < 1 strictly (no int8 overflow of slice 0)
    out = []
    for _ in range(s):
        R = R * 2.0 ** beta
        P = np.trunc(R)
        out.append(P.astype(np.int64))
        R = R - P
    return e, out


def emulate(A, B, beta, s):
    ea, As = slices(A, beta, s)
    eb, Bs = slices(B, beta, s)
    C = np.zeros((A.shape[0], B.shape[0]), dtype=np.longdouble)
    n_gemm = 0
    for p in range(s):
        for q in range(s - p):
            prod = As[p] @ Bs[q].T                     # exact integers (INT32 on the tensor core)
            assert np.max(np.abs(prod)) < 2 ** 31
            C += prod.astype(np.longdouble) * np.longdouble(2.0) ** (-beta * (p + q + 2))
            n_gemm += 1
    return C * (2.0 ** ea) * (2.0 ** eb).T, n_gemm


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 79800
    M = N = 24
    rng = np.random.default_rng(7)
    U = rng.standard_normal((M, K))
    W = rng.uniform(-5e-4, 5e-4, (N, K))
    exact = U.astype(np.longdouble) @ W.astype(np.longdouble).T
    fp64 = U @ W.T
    scale = float(np.max(np.abs(exact)))
    print(f"K = {K}, max|C| = {scale:.3e}, plain FP64 GEMM error / max|C| = "
          f"{float(np.max(np.abs(fp64 - exact))) / scale:.2e}")
    print("beta  slices  slice-GEMMs  max error / max|C|   INT32 headroom (bits)")
    for beta in (6, 7):
        for s in range(4, 10):
            C, n_gemm = emulate(U, W, beta, s)
            err = float(np.max(np.abs(C - exact))) / scale
            headroom = 31 - (2 * beta + np.log2(K))
            print(f"{beta:4d}  {s:6d}  {n_gemm:11d}  {err:18.2e}   {headroom:6.1f}")


if __name__ == "__main__":
    main()
