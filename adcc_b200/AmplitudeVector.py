"""Block vector of the ADC eigenproblem: {"ph": Tensor, "pphh": Tensor}.

Same behaviour as adcc/AmplitudeVector.py:25-165: attribute access to blocks, blockwise
arithmetic with scalars or other vectors, ``__add__`` tolerating missing blocks and dropping
blocks that sum to the integer 0, ``dot`` with one vector (float) or a list (ndarray), ``@``.
"""
import numpy as np

from . import backend as be


class AmplitudeVector(dict):
    def __init__(self, **blocks):
        super().__init__(**blocks)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, item):
        if key in self:
            self[key] = item
        else:
            raise AttributeError(key)

    @property
    def blocks(self):
        return sorted(self.keys())

    @property
    def needs_evaluation(self):
        return any(t.needs_evaluation for t in self.values())

    def evaluate(self):
        for t in self.values():
            t.evaluate()
        return self

    def _map(self, fn):
        return AmplitudeVector(**{k: fn(t) for k, t in self.items()})

    def copy(self): return self._map(lambda t: t.copy())
    def ones_like(self): return self._map(lambda t: t.ones_like())
    def empty_like(self): return self._map(lambda t: t.empty_like())
    def nosym_like(self): return self._map(lambda t: t.nosym_like())
    def zeros_like(self): return self._map(lambda t: t.zeros_like())

    def set_random(self):
        for t in self.values():
            t.set_random()
        return self

    def _dot_plan(self):
        """(weight per block, slab layout or None).  Packed doubles count every stored element 4
        times (Frobenius product of the dense tensors, adcc/AmplitudeVector.py:85-95); when the
        doubles are distributed over GPUs the replicated blocks are counted on rank 0 only and the
        partial sums are completed by one all-reduce."""
        layout = None
        for t in self.values():
            layout = getattr(t, "slab", None) or layout
        weights = []
        for b in self.keys():
            w = getattr(self[b], "dot_weight", 1.0)
            if layout is not None and getattr(self[b], "slab", None) is None and layout.rank != 0:
                w = 0.0
            weights.append(w)
        return weights, layout

    def dot(self, other):
        """Scalar product(s) summed over the blocks on the device: one host read per call."""
        single = not isinstance(other, list)
        others = [other] if single else other
        if len(others) == 0:
            return np.zeros(0)
        weights, layout = self._dot_plan()
        pairs = []
        for b in self.keys():
            for av in others:
                self[b]._check_same(av[b])
            pairs.append((self[b]._contig(), [av[b]._contig() for av in others]))
        if layout is None:
            res = be.dot_many_blocks(pairs, scales=weights)
        else:
            out = be.empty((len(others),))
            be.dot_many_blocks(pairs, scales=weights, out=out)
            be.comm_all_reduce(out, group=layout.group)          # Gram-row all-reduce
            res = be.to_host(out)
        return float(res[0]) if single else res

    def __matmul__(self, other):
        if isinstance(other, AmplitudeVector):
            return self.dot(other)
        if isinstance(other, list) and all(isinstance(e, AmplitudeVector) for e in other):
            return self.dot(other)
        return NotImplemented

    def _blockwise(self, fname, other):
        if isinstance(other, AmplitudeVector):
            if sorted(other.blocks) != sorted(self.blocks):
                raise ValueError(f"Blocks of both AmplitudeVector objects need to agree to perform {fname}")
            ret = {k: getattr(t, fname)(other[k]) for k, t in self.items()}
        else:
            ret = {k: getattr(t, fname)(other) for k, t in self.items()}
        if any(r is NotImplemented for r in ret.values()):
            return NotImplemented
        return AmplitudeVector(**ret)

    def __mul__(self, o): return self._blockwise("__mul__", o)
    def __rmul__(self, o): return self._blockwise("__rmul__", o)
    def __sub__(self, o): return self._blockwise("__sub__", o)
    def __rsub__(self, o): return self._blockwise("__rsub__", o)
    def __truediv__(self, o): return self._blockwise("__truediv__", o)
    def __imul__(self, o): return self._blockwise("__imul__", o)
    def __iadd__(self, o): return self._blockwise("__iadd__", o)
    def __isub__(self, o): return self._blockwise("__isub__", o)
    def __itruediv__(self, o): return self._blockwise("__itruediv__", o)

    def __add__(self, other):
        if isinstance(other, AmplitudeVector):
            ret = {}
            for k in sorted(set(self.blocks) | set(other.blocks)):
                a, b = self.get(k, 0), other.get(k, 0)
                v = b if isinstance(a, int) and a == 0 else (a if isinstance(b, int) and b == 0 else a + b)
                if not (isinstance(v, int) and v == 0):
                    ret[k] = v
            return AmplitudeVector(**ret)
        return AmplitudeVector(**{k: t + other for k, t in self.items()})

    def __radd__(self, other):
        if isinstance(other, AmplitudeVector):
            return other.__add__(self)
        if isinstance(other, int) and other == 0:   # sum() start value
            return self
        return AmplitudeVector(**{k: other + t for k, t in self.items()})

    def __repr__(self):
        return "AmplitudeVector(" + "=..., ".join(self.blocks) + "=...)"


def dot_rows(rows):
    """[x @ ys for x, ys in rows] with a single device-to-host read: every row's batched dot writes
    its slice of one device array (used for the Davidson subspace-matrix update)."""
    rows = [(x, list(ys)) for x, ys in rows]
    if not rows:
        return []
    total = sum(len(ys) for _, ys in rows)
    if total == 0:
        return [np.zeros(0) for _ in rows]
    out = be.empty((total,))
    pos, cuts, layout = 0, [], None
    for x, ys in rows:
        if ys:
            weights, lay = x._dot_plan()
            layout = lay or layout
            be.dot_many_blocks([(x[b]._contig(), [av[b]._contig() for av in ys]) for b in x.keys()],
                               scales=weights, out=out[pos:pos + len(ys)])
        cuts.append((pos, pos + len(ys)))
        pos += len(ys)
    if layout is not None:
        be.comm_all_reduce(out, group=layout.group)              # one all-reduce per batch of Gram rows
    host = be.to_host(out)
    return [host[a:b] for a, b in cuts]
