"""Minimal stand-in for the third-party package ``opt_einsum`` (>= 3.0), which adcc uses to split an
einsum into pairwise contractions (adcc/functions.py:198-219) and which is not installed here.
TEST INFRASTRUCTURE for running the reference's own Python layer; with the real opt_einsum on the
path this directory is not needed.

Implements what adcc relies on (opt_einsum/contract.py of the published package): ``contract`` with
a flop-greedy pairwise path, each pair evaluated with the backend module's ``tensordot`` +
``transpose`` when it is a plain contraction and with the backend's registered einsum otherwise;
``parser.gen_unused_symbols``; the ``backends.dispatch`` registries adcc fills in
(adcc/opt_einsum_integration.py:116-132)."""
import importlib

from . import parser  # noqa: F401
from .backends import dispatch as _dispatch

__version__ = "3.3.0-standin"


def _get_func(name, backend, default=None):
    key = (name, backend)
    if key in _dispatch._cached_funcs:
        return _dispatch._cached_funcs[key]
    lib = importlib.import_module(backend)
    fn = getattr(lib, name, default)
    if fn is None:
        raise AttributeError(f"backend {backend} has no function {name}")
    _dispatch._cached_funcs[key] = fn
    return fn


def _parse(subscripts, n_ops):
    subscripts = subscripts.replace(" ", "")
    if "->" in subscripts:
        lhs, out = subscripts.split("->")
    else:
        lhs = subscripts
        flat = lhs.replace(",", "")
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
    ins = lhs.split(",")
    if len(ins) != n_ops:
        raise ValueError("Number of einsum subscripts must be equal to the number of operands.")
    return ins, out


def _pair(ia, a, ib, b, keep, backend):
    """Contract two operands; ``keep``: indices needed afterwards."""
    shared = [c for c in ia if c in ib]
    out = [c for c in dict.fromkeys(ia + ib) if c in keep]
    plain = (len(set(ia)) == len(ia) and len(set(ib)) == len(ib) and not any(c in keep for c in shared))
    if plain:
        axes = ([ia.index(c) for c in shared], [ib.index(c) for c in shared])
        res = _get_func("tensordot", backend)(a, b, axes=axes)
        cur = [c for c in ia if c not in shared] + [c for c in ib if c not in shared]
        if cur != out and not isinstance(res, float):
            perm = tuple(cur.index(c) for c in out)
            res = _get_func("transpose", backend, lambda x, axes: x.transpose(axes))(res, perm)
        return "".join(out), res
    einsum = _get_func("einsum", backend)
    return "".join(out), einsum(f"{ia},{ib}->{''.join(out)}", a, b)


def contract(subscripts, *operands, optimize="auto", backend="numpy", **kwargs):
    ins, out = _parse(subscripts, len(operands))
    ops = list(operands)
    if len(ops) == 1 or any(len(set(i)) != len(i) for i in ins) and len(ops) <= 2:
        return _get_func("einsum", backend)(",".join(ins) + "->" + out, *ops)
    sizes = {}
    for idx, t in zip(ins, ops):
        for c, n in zip(idx, t.shape):
            sizes[c] = n
    while len(ops) > 1:
        best = None
        for i in range(len(ops)):
            for j in range(i + 1, len(ops)):
                rest = "".join(ins[k] for k in range(len(ops)) if k not in (i, j)) + out
                res = [c for c in dict.fromkeys(ins[i] + ins[j]) if c in rest]
                flops = 1
                for c in set(ins[i] + ins[j]):
                    flops *= sizes[c]
                mem = 1
                for c in res:
                    mem *= sizes[c]
                if best is None or (flops, mem) < best[0]:
                    best = ((flops, mem), i, j)
        _, i, j = best
        rest = "".join(ins[k] for k in range(len(ops)) if k not in (i, j)) + out
        new_idx, new = _pair(ins[i], ops[i], ins[j], ops[j], rest, backend)
        ops = [ops[k] for k in range(len(ops)) if k not in (i, j)] + [new]
        ins = [ins[k] for k in range(len(ins)) if k not in (i, j)] + [new_idx]
    res, idx = ops[0], ins[0]
    if isinstance(res, float) or idx == out:
        return res
    perm = tuple(idx.index(c) for c in out)
    return _get_func("transpose", backend, lambda x, axes: x.transpose(axes))(res, perm)
