"""Tensor algebra front-end: einsum, tensordot, direct_sum, lincomb, evaluate, dot.

Mirrors adcc/functions.py:30-219 and adcc/opt_einsum_integration.py:28-132 (diagonal /
permutation fall-backs), plus libadcc.tensordot / direct_sum / linear_combination_strict
(libadcc_src/pyiface/export_Tensor.cc:198-277).  opt_einsum is not used: a pairwise
contraction is mapped straight onto the DMMA GEMM (transpose-transpose-GEMM-transpose with
cached layouts for constant operands), multi-operand expressions are reduced pairwise by a
flop-greedy search.
"""
import numbers
import numpy as np
import torch

from . import backend as be
from .Tensor import Tensor
from .AmplitudeVector import AmplitudeVector


# ------------------------------------------------------------------------------ small helpers
def evaluate(a):
    if isinstance(a, list):
        return [evaluate(x) for x in a]
    if hasattr(a, "evaluate"):
        return a.evaluate()
    return a


def dot(a, b):
    return a.dot(b)


def copy(a): return a.copy()
def transpose(a, axes=None): return a.transpose(axes)
def empty_like(a): return a.empty_like()
def nosym_like(a): return a.nosym_like()
def zeros_like(a): return a.zeros_like()
def ones_like(a): return a.ones_like()
def contract(contraction, a, b): return einsum(contraction, a, b)


def linear_combination_strict(coefficients, tensors):
    """libadcc_src/pyiface/export_Tensor.cc:260-277 -> one fused multi-AXPY launch."""
    coefficients = np.asarray(coefficients, dtype=float)
    if len(coefficients) != len(tensors):
        raise ValueError("Number of coefficient values does not match number of tensors")
    if len(tensors) == 0:
        raise ValueError("List of tensors cannot be empty")
    first = tensors[0]
    for t in tensors[1:]:
        first._check_same(t)
    out = be.lincomb(coefficients, [t._contig() for t in tensors])
    return first._like(out)


def lincomb(coefficients, tensors, evaluate=False):
    """adcc/functions.py:92-135."""
    if len(tensors) == 0:
        raise ValueError("List of tensors cannot be empty")
    if len(tensors) != len(coefficients):
        raise ValueError("Number of coefficient values does not match number of tensors.")
    if isinstance(tensors[0], AmplitudeVector):
        return AmplitudeVector(**{
            block: lincomb(coefficients, [ten[block] for ten in tensors], evaluate=evaluate)
            for block in tensors[0].blocks
        })
    if not isinstance(tensors[0], Tensor):
        raise TypeError("Tensor type not supported")
    return linear_combination_strict(coefficients, tensors)


def lincomb_many(coefficient_rows, tensors):
    """[lincomb(row, tensors, evaluate=True) for row in coefficient_rows] in one pass over the
    inputs per block (adccb_lincomb_multi); bit-identical to the separate calls."""
    rows = np.atleast_2d(np.asarray(coefficient_rows, dtype=float))
    if len(tensors) == 0:
        raise ValueError("List of tensors cannot be empty")
    if rows.shape[1] != len(tensors):
        raise ValueError("Number of coefficient values does not match number of tensors.")
    if len(rows) == 0:
        return []
    if isinstance(tensors[0], AmplitudeVector):
        per_block = {block: lincomb_many(rows, [ten[block] for ten in tensors]) for block in tensors[0].blocks}
        return [AmplitudeVector(**{b: per_block[b][j] for b in per_block}) for j in range(len(rows))]
    if not isinstance(tensors[0], Tensor):
        raise TypeError("Tensor type not supported")
    first = tensors[0]
    for t in tensors[1:]:
        first._check_same(t)
    outs = be.lincomb_multi(rows, [t._contig() for t in tensors])
    return [first._like(o) for o in outs]


linear_combination = lincomb


# ------------------------------------------------------------------------------ contraction
def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


def _layout2d(t, idx, target, rows):
    """View of `t` (indices idx) as a 2-D matrix with index order `target`; first `rows`
    target indices form the row index.  Permutes on the device only when needed."""
    axes = tuple(idx.index(c) for c in target)
    data = t.layout(axes) if isinstance(t, Tensor) else (
        t if (axes == tuple(range(t.dim())) and t.is_contiguous()) else be.permute(t, axes))
    shape = data.shape
    return data.reshape(_prod(shape[:rows]), _prod(shape[rows:]))


def _contract_pair(ia, A, ib, B, io, alpha=1.0):
    """sum over indices shared by ia and ib that are absent from io.  A, B: Tensor objects.
    Returns (torch data with index order io)."""
    sizes = {}
    for idx, t in ((ia, A), (ib, B)):
        for c, n in zip(idx, t.shape):
            if sizes.setdefault(c, n) != n:
                raise ValueError(f"Extent mismatch for index {c}")
    shared = [c for c in ia if c in ib]
    contracted = [c for c in shared if c not in io]
    batch = [c for c in shared if c in io]
    if batch:
        if set(ia) != set(ib):
            raise NotImplementedError("batched (Hadamard) indices need identical index sets")
        order = batch + contracted
        a2 = _layout2d(A, ia, order, len(batch))
        b2 = _layout2d(B, ib, order, len(batch))
        prod = be.mul(be.contiguous(a2), be.contiguous(b2), alpha)
        ones = be.empty((1, prod.shape[1]))
        be.fill(ones, 1.0)
        res = be.gemm(prod, ones).reshape([sizes[c] for c in batch])
        if list(io) != batch:
            res = be.permute(res, [batch.index(c) for c in io])
        return res
    freeA = [c for c in ia if c not in contracted]
    freeB = [c for c in ib if c not in contracted]
    if sorted(freeA + freeB) != sorted(io):
        raise ValueError(f"Invalid contraction {ia},{ib}->{io}")

    def cost(order):
        c = 0
        if list(ia) != freeA + order:
            c += A.size
        if list(ib) != freeB + order:
            c += B.size
        return c
    orderA = [c for c in ia if c in contracted]
    orderB = [c for c in ib if c in contracted]
    order = orderA if cost(orderA) <= cost(orderB) else orderB
    a2 = _layout2d(A, ia, freeA + order, len(freeA))
    b2 = _layout2d(B, ib, freeB + order, len(freeB))
    if len(io) == 0:
        return alpha * be.dot(be.contiguous(a2), be.contiguous(b2))
    if list(io) == freeB + freeA and freeA and freeB:
        c = be.gemm(b2, a2, alpha=alpha)
        return c.reshape([sizes[x] for x in io])
    c = be.gemm(a2, b2, alpha=alpha).reshape([sizes[x] for x in freeA + freeB])
    if list(io) != freeA + freeB:
        cur = freeA + freeB
        c = be.permute(c, [cur.index(x) for x in io])
    return c


def _spaces_of(indices, operands_idx, operands):
    sp = {}
    for idx, t in zip(operands_idx, operands):
        for c, s in zip(idx, t.subspaces):
            if sp.setdefault(c, s) != s:
                raise ValueError(f"Index {c} refers to different subspaces ({sp[c]} vs {s}); only "
                                 "equivalent axes may be contracted (TensorImpl.cc:934-940)")
    return [sp[c] for c in indices]


def _single_operand(idx, t, out):
    """diagonal extraction / permutation / full trace of one tensor
    (adcc/opt_einsum_integration.py:28-113)."""
    for c in set(idx):
        axes = [k for k, x in enumerate(idx) if x == c]
        if len({t.subspaces[a] for a in axes}) != 1:
            raise ValueError("repeated index over different subspaces")
    summed = [c for c in set(idx) if c not in out]
    data = t._data
    if summed:
        if len(out) != 0:
            raise NotImplementedError("Partial traces are not supported (opt_einsum_integration.py:68-71)")
        letters = sorted(set(idx), key=idx.index)
        shape = [data.shape[idx.index(c)] for c in letters]
        strides = [sum(data.stride(k) for k, x in enumerate(idx) if x == c) for c in letters]
        diag = be.gather(data, shape, strides)
        ones = torch.empty_like(diag)
        be.fill(ones, 1.0)
        return be.dot(diag, ones)
    shape = [data.shape[idx.index(c)] for c in out]
    strides = [sum(data.stride(k) for k, x in enumerate(idx) if x == c) for c in out]
    res = be.gather(data, shape, strides)
    return Tensor._wrap(t.mospaces, [t.subspaces[idx.index(c)] for c in out], res)


def einsum(subscripts, *operands, optimise="auto"):
    """Einstein summation over device tensors (adcc/functions.py:198-219)."""
    subscripts = subscripts.replace(" ", "")
    if "->" in subscripts:
        lhs, out = subscripts.split("->")
    else:
        lhs = subscripts
        flat = lhs.replace(",", "")
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
    ins = lhs.split(",")
    if len(ins) != len(operands):
        raise ValueError("Number of subscripts does not match the number of operands")
    for idx, t in zip(ins, operands):
        if not isinstance(t, Tensor):
            raise TypeError("einsum operands need to be Tensors")
        if len(idx) != t.ndim:
            raise ValueError(f"Subscript {idx} does not match tensor rank {t.ndim}")
    ops = list(operands)
    ins = [str(i) for i in ins]
    # repeated indices inside one operand -> take the diagonal first
    for k, (idx, t) in enumerate(zip(ins, ops)):
        if len(set(idx)) != len(idx):
            if len(ops) == 1:
                return _single_operand(idx, t, out)
            keep = "".join(sorted(set(idx), key=idx.index))
            ops[k] = _single_operand(idx, t, keep)
            ins[k] = keep
    if len(ops) == 1:
        return _single_operand(ins[0], ops[0], out)
    mosp = ops[0].mospaces
    while len(ops) > 2:
        # flop-greedy choice of the next pair
        best = None
        for i in range(len(ops)):
            for j in range(i + 1, len(ops)):
                others = "".join(ins[k] for k in range(len(ops)) if k not in (i, j)) + out
                res = "".join(c for c in dict.fromkeys(ins[i] + ins[j]) if c in others)
                sizes = {c: n for idx, t in ((ins[i], ops[i]), (ins[j], ops[j])) for c, n in zip(idx, t.shape)}
                flops = _prod(sizes.values())
                mem = _prod(sizes[c] for c in res)
                key = (flops, mem)
                if best is None or key < best[0]:
                    best = (key, i, j, res)
        _, i, j, res = best
        sp = _spaces_of(res, [ins[i], ins[j]], [ops[i], ops[j]])
        data = _contract_pair(ins[i], ops[i], ins[j], ops[j], res)
        new = Tensor._wrap(mosp, sp, data)
        ops = [ops[k] for k in range(len(ops)) if k not in (i, j)] + [new]
        ins = [ins[k] for k in range(len(ins)) if k not in (i, j)] + [res]
    sp = _spaces_of(out, ins, ops)
    data = _contract_pair(ins[0], ops[0], ins[1], ops[1], out)
    if len(out) == 0:
        return float(data)
    return Tensor._wrap(mosp, sp, data)


def tensordot(a, b, axes=2):
    """libadcc.tensordot (export_Tensor.cc:198-218): numpy.tensordot semantics."""
    if isinstance(axes, numbers.Integral):
        ax_a = list(range(a.ndim - axes, a.ndim))
        ax_b = list(range(axes))
    else:
        ax_a, ax_b = axes
        ax_a = [ax_a] if isinstance(ax_a, numbers.Integral) else list(ax_a)
        ax_b = [ax_b] if isinstance(ax_b, numbers.Integral) else list(ax_b)
    if len(ax_a) != len(ax_b):
        raise ValueError("Length of the passed axes does not agree")
    letters = "abcdefghijklmnopqrstuvwxyz"
    ia = list(letters[:a.ndim])
    ib = list(letters[a.ndim:a.ndim + b.ndim])
    for x, y in zip(ax_a, ax_b):
        ib[y] = ia[x]
    out = [c for k, c in enumerate(ia) if k not in ax_a] + [c for k, c in enumerate(ib) if k not in ax_b]
    return einsum("".join(ia) + "," + "".join(ib) + "->" + "".join(out), a, b)


def trace(subscripts, tensor):
    return einsum(subscripts + "->", tensor)


def direct_sum(subscripts, *operands):
    """adcc/functions.py:156-195: e.g. direct_sum("a-i->ia", fvv.diagonal(), foo.diagonal())."""
    subscripts = subscripts.replace(" ", "")

    def split_signs_symbols(s):
        s = s.replace(",", "+")
        if s[0] not in "+-":
            s = "+" + s
        signs = [x for x in s if x in "+-"]
        symbols = s[1:].replace("-", "+").split("+")
        return signs, symbols

    if "->" in subscripts:
        src, dest = subscripts.split("->")
        signs, src = split_signs_symbols(src)
        permutation = tuple("".join(src).index(c) for c in dest)
    else:
        signs, src = split_signs_symbols(subscripts)
        permutation = None
    if len(src) != len(operands):
        raise ValueError("Number of contraction subscripts does not agree with number of operands")
    for i, idcs in enumerate(src):
        if len(idcs) != operands[i].ndim:
            raise ValueError(f"Number of subscripts of {i}-th tensor (== {idcs}) does not match "
                             f"dimension of tensor (== {operands[i].ndim}).")
    res = operands[0]._contig()
    sp = list(operands[0].subspaces)
    sa = -1.0 if signs[0] == "-" else 1.0
    if len(operands) == 1:
        res = be.scaled_copy(res, sa)
    for i, op in enumerate(operands[1:]):
        sb = -1.0 if signs[i + 1] == "-" else 1.0
        res = be.direct_sum(res, op._contig(), sa, sb)
        sa = 1.0
        sp = sp + list(op.subspaces)
    out = Tensor._wrap(operands[0].mospaces, sp, res)
    if permutation is not None:
        out = out.transpose(permutation)
        out._contig()
    return out
