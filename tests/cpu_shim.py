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
    def axpby(a, x, b, y):
        y.copy_(a * x.reshape(y.shape) + (b * y if b != 0.0 else 0.0))   # beta == 0 never reads y
        return y
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

    def lincomb_multi(coeff_rows, tensors):
        return [lincomb(row, tensors) for row in np.asarray(coeff_rows, dtype=float)]

    def dot_many(x, ys, scale=1.0):
        return np.array([scale * float(torch.dot(x.reshape(-1), y.reshape(-1))) for y in ys])

    def dot(x, y): return float(torch.dot(x.reshape(-1), y.reshape(-1)))

    def dot_many_blocks(pairs, scale=1.0, out=None, scales=None):
        res = sum(dot_many(x, ys, scale if scales is None else scales[ib]) for ib, (x, ys) in enumerate(pairs))
        if out is None:
            return res
        out.copy_(torch.from_numpy(np.asarray(res)))
        return out

    def to_host(t): return t.numpy()
    def upload(t): return t.clone()
    def download(dev, host): host.copy_(dev); return host

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


    def precond_apply(x, D, shift, n_alpha, sym_mode=0, spin_fac=0.0, purify=False):
        p = x / (D - shift) if D is not None else x.clone()
        if spin_fac != 0.0:
            p = spin_project(p, n_alpha, spin_fac)
        if x.dim() == 4 and sym_mode == 1:
            p = 0.5 * (p - p.permute(1, 0, 2, 3))
            p = 0.5 * (p - p.permute(0, 1, 3, 2))
            p = 0.5 * (p + p.permute(1, 0, 3, 2))
        elif x.dim() == 4 and sym_mode == 2:
            p = 0.5 * (p - p.permute(0, 1, 3, 2))
        p = p.contiguous()
        if x.dim() == 4 and purify:
            spin_singlet_(p, n_alpha)
        return p

    # ---- packed doubles primitives (numpy restatements of csrc/packed_ops.cu)
    def _pairs(n):
        return [(p, q) for q in range(n) for p in range(q)]

    def packed_unpack(U, no, nv):
        out = torch.zeros((no, no, nv, nv), dtype=F64)
        po, pv = _pairs(no), _pairs(nv)
        Un = U.numpy()
        o = out.numpy()
        for I, (i, j) in enumerate(po):
            for A, (a, b) in enumerate(pv):
                x = Un[I, A]
                o[i, j, a, b] = x; o[j, i, a, b] = -x; o[i, j, b, a] = -x; o[j, i, b, a] = x
        return out

    def packed_pack(T, no, nv):
        t = T.contiguous().numpy().reshape(no, no, nv, nv)
        t = 0.25 * (t - t.transpose(1, 0, 2, 3) - t.transpose(0, 1, 3, 2) + t.transpose(1, 0, 3, 2))
        po, pv = _pairs(no), _pairs(nv)
        ii = np.array([p[0] for p in po], dtype=int); jj = np.array([p[1] for p in po], dtype=int)
        aa = np.array([p[0] for p in pv], dtype=int); bb = np.array([p[1] for p in pv], dtype=int)
        if len(po) == 0 or len(pv) == 0:
            return torch.zeros((len(po), len(pv)), dtype=F64)
        return from_numpy(t[ii[:, None], jj[:, None], aa[None, :], bb[None, :]])

    def packed_expand(which, U, no, nv, out=None, n_pairs_o=None):
        res = _packed_expand(which, U, no, nv, n_pairs_o)
        if out is not None:
            out.copy_(res.reshape(out.shape))
            return out
        return res

    def _packed_expand(which, U, no, nv, n_pairs_o=None):
        if n_pairs_o is not None:
            assert which == 2
            pv = _pairs(nv)
            res = torch.zeros((nv, n_pairs_o, nv), dtype=F64)
            for A, (a, b) in enumerate(pv):
                res[a, :, b] = U[:, A]
                res[b, :, a] = -U[:, A]
            return res.reshape(nv, -1)
        d = packed_unpack(U, no, nv)
        if which == 0:
            return d.permute(0, 2, 1, 3).contiguous().reshape(no * nv, no * nv)
        if which == 1:
            pv = _pairs(nv)
            aa = torch.tensor([p[0] for p in pv], dtype=torch.long); bb = torch.tensor([p[1] for p in pv], dtype=torch.long)
            return d[:, :, aa, bb].contiguous().reshape(no, -1)
        po = _pairs(no)
        ii = torch.tensor([p[0] for p in po], dtype=torch.long); jj = torch.tensor([p[1] for p in po], dtype=torch.long)
        return d[ii, jj].permute(1, 0, 2).contiguous().reshape(nv, -1)

    def packed_expand_slab(which, U, no, nv, first, count):
        if which == 2:
            return packed_expand(2, U, no, nv, n_pairs_o=count)
        npv = nv * (nv - 1) // 2
        full = packed_expand(which, U, no, nv)
        if which == 0:
            return full.reshape(no, nv, no * nv)[first:first + count].reshape(count * nv, no * nv).contiguous()
        return full.reshape(no, no, npv)[:, first:first + count].reshape(no, count * npv).contiguous()

    def slab_layout(direction, full, slabs, cuts, wmax):
        for r in range(len(cuts) - 1):
            w = cuts[r + 1] - cuts[r]
            if direction == 0:
                slabs[r].zero_()
                slabs[r][:, :w] = full[:, cuts[r]:cuts[r + 1]]
            else:
                full[:, cuts[r]:cuts[r + 1]] = slabs[r][:, :w]
        return slabs if direction == 0 else full

    def copy2d(dst, src, nrows, ncols):
        dst[:nrows, :ncols] = src[:nrows, :ncols]
        return dst

    def comm_reduce_scatter(out, inp, group=None):
        import torch.distributed as dist          # gloo has no reduce-scatter: all-reduce + own chunk
        tmp = inp.clone()
        dist.all_reduce(tmp, group=group)
        n = out.numel()
        r = dist.get_rank(group)
        out.copy_(tmp.reshape(-1)[r * n:(r + 1) * n].reshape(out.shape))
        return out

    def comm_all_gather(out, inp, group=None):
        import torch.distributed as dist
        parts = [torch.empty_like(inp) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, inp.contiguous(), group=group)
        out.copy_(torch.stack(parts).reshape(out.shape))
        return out

    def block_gather(inp, rows=None, cols=None):
        res = inp if rows is None else inp[rows]
        return (res if cols is None else res[:, cols]).contiguous()

    def block_scatter_add(out, blk, rows=None, cols=None, alpha=1.0, assign=False):
        r = torch.arange(out.shape[0]) if rows is None else rows
        c = torch.arange(out.shape[1]) if cols is None else cols
        if assign:
            out[r[:, None], c[None, :]] = alpha * blk
        else:
            out[r[:, None], c[None, :]] += alpha * blk
        return out

    def _pair_dense(rows, nv):
        """[n, npv] pair-packed antisymmetric rows -> [n, nv, nv]"""
        pv = _pairs(nv)
        aa = torch.tensor([p[0] for p in pv], dtype=torch.long); bb = torch.tensor([p[1] for p in pv], dtype=torch.long)
        d = torch.zeros((rows.shape[0], nv, nv), dtype=F64)
        d[:, aa, bb] = rows
        d[:, bb, aa] = -rows
        return d

    def vvvv_exchange_slab(W, nv, a_begin, na):
        npv = nv * (nv - 1) // 2
        full = _pair_dense(_pair_dense(W.reshape(npv, npv), nv).reshape(npv, nv * nv).t().contiguous(), nv)
        full = full.reshape(nv, nv, nv, nv)           # [c', d', a', b'] = <a'b'||c'd'> -> index as [b,d,a,c]
        # <ac||bd> = full[b, d, a, c]
        return full.permute(2, 0, 3, 1)[a_begin:a_begin + na].contiguous()

    def ovvv_exchange_slab(V2, no, nv, a_begin, na):
        npv = nv * (nv - 1) // 2
        d = _pair_dense(V2.reshape(nv * no, npv), nv).reshape(nv, no, nv, nv)      # [c, j, a, d]
        return d.permute(2, 0, 1, 3)[a_begin:a_begin + na].contiguous()

    def pair_matvec(V, nv, n_out, n_red, row_stride_out, row_stride_red, X, x_by_out, out=None, alpha=1.0, beta=0.0):
        npv = nv * (nv - 1) // 2
        rows = V.reshape(-1, npv)
        res = torch.zeros((n_out, nv), dtype=F64)
        for o in range(n_out):
            idx = torch.tensor([o * row_stride_out + r * row_stride_red for r in range(n_red)], dtype=torch.long)
            A = _pair_dense(rows[idx], nv)
            xs = X[o].reshape(1, nv).expand(n_red, nv) if x_by_out else X[:n_red]
            res[o] = torch.einsum("rac,rc->a", A, xs)
        if out is None:
            return alpha * res
        out.copy_(alpha * res + (beta * out if beta != 0.0 else 0.0))
        return out

    def packed_fold_virt(S, nv, sa, P1, off1, P2=None, off2=None, alpha=1.0, rows=None):
        pv = _pairs(nv)
        aa = np.array([p[0] for p in pv], dtype=int); bb = np.array([p[1] for p in pv], dtype=int)
        p1 = P1.contiguous().reshape(-1).numpy()
        p2 = P2.contiguous().reshape(-1).numpy() if P2 is not None else None
        Sn = S.numpy()
        for r in range(len(off1)):
            def F(a, b):
                v = p1[int(off1[r]) + a * sa + b]
                if p2 is not None:
                    v = v - p2[int(off2[r]) + a * sa + b]
                return v
            Sn[int(rows[r]) if rows is not None else r] += alpha * (F(aa, bb) - F(bb, aa))
        return S

    def packed_fold_occ(S, C, no, nv, alpha=1.0, j_begin=0, nj=None):
        nj = no if nj is None else nj
        c = C.contiguous().reshape(no, nj, -1)
        for i in range(no):
            for jl in range(nj):
                j = j_begin + jl
                if i == j:
                    continue
                row = max(i, j) * (max(i, j) - 1) // 2 + min(i, j)
                S[row] += (alpha if i < j else -alpha) * c[i, jl]
        return S

    def packed_diagonal(no, nv, dov, doo=None, dvv=None):
        po, pv = _pairs(no), _pairs(nv)
        D = torch.zeros((len(po), len(pv)), dtype=F64)
        dov = dov.reshape(no, nv)
        for I, (i, j) in enumerate(po):
            for A, (a, b) in enumerate(pv):
                v = 0.5 * (dov[i, a] + dov[j, b] + dov[i, b] + dov[j, a])
                if doo is not None:
                    v = v + doo.reshape(no, no)[i, j]
                if dvv is not None:
                    v = v + dvv.reshape(nv, nv)[a, b]
                D[I, A] = v
        return D

    def pair_select(t, axis):
        n = t.shape[axis]
        pp = _pairs(n)
        ii = torch.tensor([p[0] for p in pp], dtype=torch.long); jj = torch.tensor([p[1] for p in pp], dtype=torch.long)
        idx = [slice(None)] * t.dim()
        idx[axis] = ii
        idx[axis + 1] = jj
        res = t.contiguous()[tuple(idx)]
        # advanced indices adjacent -> result axis stays in place
        return res.contiguous()

    def add_columns(S, col0, G, ncols):
        S[:, col0:col0 + ncols] += G[:, :ncols]
        return S

    def spin_expand4(T):
        n = T.shape
        out = torch.zeros(tuple(2 * x for x in n), dtype=F64)
        for a in (0, 1):
            for b in (0, 1):
                out[a * n[0]:(a + 1) * n[0], a * n[1]:(a + 1) * n[1], b * n[2]:(b + 1) * n[2], b * n[3]:(b + 1) * n[3]] = T
        return out

    def int64_tensor(values):
        return torch.tensor([int(v) for v in values], dtype=torch.int64)

    def synth_fill(out, layout, no, nv, seed, scale, block="vvvv", row_begin=0, j_begin=0, nj=0):
        nj = no if nj <= 0 else nj
        osl = np.arange(j_begin, j_begin + nj)
        from adcc_b200.synthetic import synthetic_eri_value as val
        po, pv = np.array(_pairs(no), dtype=int).reshape(-1, 2), np.array(_pairs(nv), dtype=int).reshape(-1, 2)
        o, v = np.arange(no), np.arange(nv)
        if layout == "vvvv_packed":
            rows = pv[row_begin:row_begin + out.shape[0]]
            arr = val(seed, "vvvv", rows[:, 0][:, None], rows[:, 1][:, None], pv[:, 0][None, :], pv[:, 1][None, :], scale)
        elif layout == "oooo_packed":
            arr = val(seed, "oooo", po[:, 0][:, None], po[:, 1][:, None], po[:, 0][None, :], po[:, 1][None, :], scale)
        elif layout == "ovov_jbkc":
            j, b, k, c = np.meshgrid(osl, v, o, v, indexing="ij")
            arr = val(seed, "ovov", k, b, j, c, scale)
        elif layout == "ovvv_jabc":
            j, AB, c = np.meshgrid(osl, np.arange(len(pv)), v, indexing="ij")
            arr = val(seed, "ovvv", j, c, pv[AB, 0], pv[AB, 1], scale)
        elif layout == "ovvv_ajbc":
            a, j, BC = np.meshgrid(v, osl, np.arange(len(pv)), indexing="ij")
            arr = val(seed, "ovvv", j, a, pv[BC, 0], pv[BC, 1], scale)
        elif layout == "ooov_ijkb":
            i, JK, b = np.meshgrid(o, np.arange(len(po)), v, indexing="ij")
            arr = val(seed, "ooov", po[JK, 0], po[JK, 1], i, b, scale)
        elif layout == "ooov_ijak":
            IJ, a, k = np.meshgrid(np.arange(len(po)), v, o, indexing="ij")
            arr = val(seed, "ooov", po[IJ, 0], po[IJ, 1], k, a, scale)
        elif layout == "vvvv_exchange":
            a, b, c, d = np.meshgrid(np.arange(row_begin, row_begin + out.shape[0]), v, v, v, indexing="ij")
            arr = val(seed, "vvvv", a, c, b, d, scale)
        elif layout == "ovvv_exchange":
            a, c, j, d = np.meshgrid(np.arange(row_begin, row_begin + out.shape[0]), v, o, v, indexing="ij")
            arr = val(seed, "ovvv", j, c, a, d, scale)
        else:
            from adcc_b200.synthetic import synthetic_eri_block
            arr = synthetic_eri_block(seed, block, no, nv, scale)
        out.copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).reshape(out.shape))
        return out

    def mul(x, y, alpha=1.0, out=None, beta=0.0):
        res = alpha * x * y
        if out is None:
            return res.contiguous()
        out.copy_(res + beta * out if beta != 0.0 else res)
        return out

    for name, fn in dict(empty=empty, zeros=zeros, from_numpy=from_numpy, gemm=gemm, gather=gather,
                         permute=permute, contiguous=contiguous, fill=fill, axpby=axpby,
                         scaled_copy=scaled_copy, mul=mul, div=div, div_shift=div_shift,
                         add_scalar=add_scalar, direct_sum=direct_sum, lincomb=lincomb, lincomb_multi=lincomb_multi,
                         dot_many=dot_many, dot=dot, dot_many_blocks=dot_many_blocks, to_host=to_host, upload=upload, download=download, spin_project=spin_project,
                         spin_singlet_=spin_singlet_, precond_apply=precond_apply, packed_unpack=packed_unpack, packed_pack=packed_pack,
                         packed_expand=packed_expand, packed_fold_virt=packed_fold_virt,
                         packed_fold_occ=packed_fold_occ, packed_diagonal=packed_diagonal,
                         pair_select=pair_select, block_gather=block_gather, block_scatter_add=block_scatter_add, vvvv_exchange_slab=vvvv_exchange_slab, ovvv_exchange_slab=ovvv_exchange_slab, pair_matvec=pair_matvec,
                         packed_expand_slab=packed_expand_slab, slab_layout=slab_layout,
                         copy2d=copy2d, comm_reduce_scatter=comm_reduce_scatter, comm_all_gather=comm_all_gather,
                         int64_tensor=int64_tensor, add_columns=add_columns, spin_expand4=spin_expand4, synth_fill=synth_fill).items():
        monkeypatch.setattr(be, name, fn)
    monkeypatch.setattr(be, "device", lambda: cpu)
    monkeypatch.setattr(be, "release_cached_memory", lambda: None)
    monkeypatch.setattr(be, "free_memory", lambda: 8 << 30)

    class ReplayGraph:
        """Host-logic stand-in: same static-buffer protocol, the launch sequence re-runs eagerly."""
        def __init__(self, shapes, fill, fn):
            self.inputs = {k: empty(shape) for k, shape in shapes.items()}
            self.fn = fn
            fill(self.inputs)
            fn(self.inputs)

        def run(self, fill):
            fill(self.inputs)
            self.outputs = self.fn(self.inputs)
            return self.outputs

    import os
    monkeypatch.setattr(be, "ReplayGraph", ReplayGraph)
    monkeypatch.setattr(be, "graphs_supported", lambda: os.environ.get("ADCCB_SHIM_GRAPHS") == "1")
