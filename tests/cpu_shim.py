"""TEST-ONLY stand-in for adcc_b200.backend (the CUDA primitives) on CPU tensors.

Lets the `-m "not gpu"` suite exercise the HOST logic (einsum planner, block tables, Davidson,
guesses, workflow) in the build container, which has no GPU.  It lives under tests/, is
installed by monkeypatching inside a fixture, and is never importable from the product: the
product's adcc_b200.backend has no CPU path and raises without the CUDA library.
"""
import numpy as np
import torch

F64 = torch.float64


def install(monkeypatch):
    from adcc_b200 import backend as be
    cpu = torch.device("cpu")

    def empty(shape): return torch.empty(tuple(int(s) for s in shape), dtype=F64)
    def zeros(shape): return torch.zeros(tuple(int(s) for s in shape), dtype=F64)
    def from_numpy(a): return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())

    def gemm(A, B, out=None, alpha=1.0, beta=0.0, tile_hint=0):
        res = alpha * (A @ B.t())
        if out is None:
            return res.contiguous()
        out.copy_(res + beta * out if beta != 0.0 else res)
        return out

    def gather(inp, shape, in_strides, alpha=1.0, add=None, gamma=0.0, out=None, beta=0.0):
        shape = [int(s) for s in shape]
        view = torch.as_strided(inp, shape, [int(s) for s in in_strides], inp.storage_offset())
        res = alpha * view
        if add is not None:
            res = res + gamma * add
        if out is not None and beta != 0.0:
            res = res + beta * out
        res = res.contiguous()
        if out is not None:
            out.copy_(res)
            return out
        return res

    def permute(t, perm, alpha=1.0, add=None, gamma=0.0, out=None, beta=0.0):
        return gather(t, [t.shape[p] for p in perm], [t.stride(p) for p in perm], alpha, add, gamma, out, beta)

    def contiguous(t): return t.contiguous()
    def fill(t, v): t.fill_(float(v)); return t
    def axpby(a, x, b, y): y.copy_(a * x.reshape(y.shape) + b * y); return y
    def scaled_copy(x, alpha=1.0): return (alpha * x).contiguous()
    def mul(x, y, alpha=1.0): return (alpha * x * y).contiguous()
    def div(x, y, alpha=1.0): return (alpha * x / y).contiguous()
    def div_shift(x, d, s): return (x / (d - s)).contiguous()
    def add_scalar(x, c): return (x + c).contiguous()

    def direct_sum(a, b, sa=1.0, sb=1.0):
        return (sa * a.reshape(tuple(a.shape) + (1,) * b.dim()) + sb * b).contiguous()

    def lincomb(coeffs, tensors, out=None, beta=0.0):
        res = sum(float(c) * t for c, t in zip(coeffs, tensors))
        if out is not None:
            out.copy_(res + beta * out if beta != 0.0 else res)
            return out
        return res.contiguous()

    def dot_many(x, ys, scale=1.0):
        return np.array([scale * float(torch.dot(x.reshape(-1), y.reshape(-1))) for y in ys])

    def dot(x, y): return float(torch.dot(x.reshape(-1), y.reshape(-1)))

    def spin_project(t, n_alpha, fac):
        f = t
        for ax, na in enumerate(n_alpha):
            f = torch.cat((f.narrow(ax, na, f.shape[ax] - na), f.narrow(ax, 0, na)), dim=ax)
        res = 0.5 * (t + fac * f)
        nd = t.dim()
        s = [torch.tensor([0] * n + [1] * (t.shape[k] - n)) for k, n in enumerate(n_alpha)]
        if nd == 2:
            keep = s[0][:, None] == s[1][None, :]
        else:
            keep = (s[0][:, None, None, None] + s[1][None, :, None, None]
                    == s[2][None, None, :, None] + s[3][None, None, None, :])
        return (res * keep).contiguous()

    def spin_singlet_(t, n_alpha):
        ni, nj, na, nb = n_alpha
        new = t[:ni, nj:, :na, nb:] + t[:ni, nj:, na:, :nb]
        t[:ni, :nj, :na, :nb] = new
        t[ni:, nj:, na:, nb:] = new
        return t

    for name, fn in dict(empty=empty, zeros=zeros, from_numpy=from_numpy, gemm=gemm, gather=gather,
                         permute=permute, contiguous=contiguous, fill=fill, axpby=axpby,
                         scaled_copy=scaled_copy, mul=mul, div=div, div_shift=div_shift,
                         add_scalar=add_scalar, direct_sum=direct_sum, lincomb=lincomb,
                         dot_many=dot_many, dot=dot, spin_project=spin_project,
                         spin_singlet_=spin_singlet_).items():
        monkeypatch.setattr(be, name, fn)
    monkeypatch.setattr(be, "device", lambda: cpu)
