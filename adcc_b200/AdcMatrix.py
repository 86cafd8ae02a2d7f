"""ADC secular matrix: block table interpreter, diagonal, matvec.

API of adcc/AdcMatrix.py:205-472 (``AdcMatrix(method, hf_or_mp, block_orders, intermediates,
diagonal_precomputed)``, ``matvec`` / ``@`` / ``rmatvec``, ``block_apply``, ``diagonal``,
``blocks``, ``axis_blocks``/``axis_spaces``/``axis_lengths``/``shape``,
``construct_symmetrisation_for_blocks``, ``block_view``) and the block-order rules of
:74-130.  The blocks are the contraction tables of adcc_b200/adc_pp.py; ``matvec`` evaluates
all terms of all blocks and merges the terms of each output block in one fused lincomb launch.
"""
import contextlib
import os
from collections import namedtuple

import numpy as np

from . import adc_pp
from .AdcMethod import AdcMethod
from .AmplitudeVector import AmplitudeVector
from .functions import einsum, linear_combination_strict
from .Intermediates import Intermediates
from .LazyMp import LazyMp
from .ReferenceState import ReferenceState
from .Tensor import Tensor
from .timings import Timer

_SPECIAL_ORDERS = {"adc2x": {"ph_ph": 2, "ph_pphh": 1, "pphh_ph": 1, "pphh_pphh": 1}}


AdcBlock = namedtuple("AdcBlock", ["apply", "diagonal"])      # adcc/adc_pp/matrix.py:41


class AdcExtraTerm:
    """Custom terms added to the blocks of an existing AdcMatrix (adcc/AdcMatrix.py:36-65): the
    plug-in point environment couplings use.  ``blocks``: {"ph_ph": callable(hf, mp, intermediates)
    -> AdcBlock(apply, diagonal)}."""

    def __init__(self, matrix, blocks):
        self.ground_state = matrix.ground_state
        self.reference_state = matrix.reference_state
        self.intermediates = matrix.intermediates
        self.blocks = {}
        if not isinstance(blocks, dict):
            raise TypeError("blocks needs to be a dict.")
        for space, block_fun in blocks.items():
            if not callable(block_fun):
                raise TypeError("Items in additional_blocks must be callable.")
            self.blocks[space] = block_fun(self.reference_state, self.ground_state, self.intermediates)


class AdcMatrixlike:
    """Marker base class for objects that behave like an ADC matrix."""


def default_block_orders(method):
    base = method.base_method.name
    if base in _SPECIAL_ORDERS:
        return dict(_SPECIAL_ORDERS[base])
    level = method.level
    spaces = ["ph", "pphh"][:level // 2 + 1]
    ret = {}
    for i1, bra in enumerate(spaces):
        for i2, ket in enumerate(spaces):
            order = level - i1 - i2
            ret[f"{bra}_{ket}"] = None if order < 0 else order
    return ret


class AdcMatrix(AdcMatrixlike):
    def __init__(self, method, hf_or_mp, block_orders=None, intermediates=None,
                 diagonal_precomputed=None, doubles_storage="auto", shard=None, vectors=None):
        """``shard=(rank, world)``: this process holds rank's slabs of the big operands (one process
        per GPU, torch.distributed initialised); ``vectors``: "slab" (default, doubles vectors
        distributed by virtual-pair slabs, SlabDoubles) or "replicated"."""
        if isinstance(hf_or_mp, ReferenceState):
            hf_or_mp = LazyMp(hf_or_mp)
        elif not isinstance(hf_or_mp, LazyMp):
            try:
                hf_or_mp = LazyMp(ReferenceState(hf_or_mp))
            except Exception:
                raise TypeError("hf_or_mp is not a valid object. It needs to be either a LazyMp, "
                                "a ReferenceState or SCF data.")
        if not isinstance(method, AdcMethod):
            method = AdcMethod(method)
        if diagonal_precomputed is not None and not isinstance(diagonal_precomputed, AmplitudeVector):
            raise TypeError("diagonal_precomputed needs to be an AmplitudeVector.")

        self.timer = Timer()
        self.method = method
        self.ground_state = hf_or_mp
        self.reference_state = hf_or_mp.reference_state
        self.mospaces = self.reference_state.mospaces
        self.is_core_valence_separated = method.is_core_valence_separated
        self.ndim = 2
        self.extra_terms = []
        self.intermediates = intermediates if intermediates is not None else Intermediates(hf_or_mp)

        hf = self.reference_state
        if hf.has_core_occupied_space and not self.is_core_valence_separated:
            raise ValueError("Cannot run a general (non-core-valence approximated) ADC method on "
                             "top of a ground state with a core-valence separation.")
        if not hf.has_core_occupied_space and self.is_core_valence_separated:
            raise ValueError("Cannot run a core-valence approximated ADC method on top of a "
                             "ground state without a core-valence separation.")

        self.block_orders = default_block_orders(method)
        if block_orders is not None:
            self.block_orders.update(block_orders)
        self._validate_block_orders()
        self._variant = "cvs" if self.is_core_valence_separated else ""
        self._const = {}
        self._graphs = {}
        self._graph_seen = {}
        self.doubles_storage = self._select_storage(doubles_storage)
        if shard is not None and (shard[1] > 1) and self.doubles_storage != "packed":
            raise NotImplementedError("multi-GPU sharding is implemented for the packed doubles store")
        self._engine = None

        with self.timer.record("build"):
            self._tables = {}
            for blk, order in self.block_orders.items():
                if order is None:
                    continue
                key = (self._variant, blk, order)
                if key not in adc_pp.BLOCKS:
                    raise ValueError(f"Could not dispatch: spaces={blk.split('_')} order={order} "
                                     f"variant={self._variant}. Probably the secular matrix is not "
                                     "implemented for the requested method.")
                self._tables[blk] = adc_pp.BLOCKS[key]
            self.blocks = {blk: self._make_apply(blk) for blk in self._tables}
            if self.doubles_storage == "packed":
                if self.is_core_valence_separated:
                    from .packed_cvs import PackedCvsEngine
                    self._engine = PackedCvsEngine(self, shard=shard, vectors=vectors)
                else:
                    from .packed import PackedAdcEngine
                    self._engine = PackedAdcEngine(self, shard=shard, vectors=vectors)
                self.blocks = {blk: self._make_engine_apply(blk) for blk in self._tables
                               if self._tables[blk][0] is not None}
                if getattr(self._engine, "diagonal_ph", None) is not None:
                    dph = AmplitudeVector(ph=self._engine.diagonal_ph)
                else:
                    dph = adc_pp.diagonal(self._variant, "ph_ph", self.block_orders["ph_ph"], hf,
                                          self.ground_state, self.intermediates)
                dph.ph._contig()
                self._diagonal = AmplitudeVector(ph=dph.ph, pphh=self._engine.diagonal_pphh)
            else:
                # resolve constant operands now: triggers ERI staging and intermediate builds
                for blk in self._tables:
                    for _, expr in self._tables[blk][2]:
                        self._resolve_constants(expr)
            if diagonal_precomputed is not None:
                self._diagonal = diagonal_precomputed
            elif self.doubles_storage == "packed":
                pass
            else:
                diag = 0
                for blk, order in self.block_orders.items():
                    if order is None:
                        continue
                    d = adc_pp.diagonal(self._variant, blk, order, hf, self.ground_state,
                                        self.intermediates)
                    if isinstance(d, AmplitudeVector):
                        diag = d if isinstance(diag, int) else diag + d
                self._diagonal = diag.evaluate()
                for t in self._diagonal.values():
                    t._contig()
            self._init_space_data(self._diagonal)

    # ------------------------------------------------------------------ bookkeeping
    def _select_storage(self, requested):
        """'dense' (generic table interpreter) or 'packed' (antisymmetry-packed doubles,
        adcc_b200/packed.py).  'auto' packs when the dense doubles vector would exceed 1 GiB or the
        vvvv term dominates (first-order pphh_pphh block, >= 96 virtual spin orbitals)."""
        if requested not in ("auto", "dense", "packed"):
            raise ValueError("doubles_storage needs to be 'auto', 'dense' or 'packed'")
        orders = tuple(self.block_orders.get(b) for b in ("ph_ph", "ph_pphh", "pphh_ph", "pphh_pphh"))
        if self.is_core_valence_separated:
            packable = orders in ((2, 1, 1, 0), (2, 1, 1, 1))       # packed_cvs.py: virtual pairs only
        else:
            packable = orders in ((2, 1, 1, 0), (2, 1, 1, 1), (3, 2, 2, 1))
        if requested == "packed" and not packable:
            raise NotImplementedError("packed doubles storage is implemented for ADC(2), ADC(2)-x, "
                                      "ADC(3), CVS-ADC(2) and CVS-ADC(2)-x")
        if requested == "auto":
            # packed: 4x less doubles / vvvv storage and 4x fewer vvvv flops; it pays once the
            # vector is large or the vvvv term dominates (measured: methyloxirane cc-pVDZ ADC(3),
            # nv = 140: 3.1 -> 1.4-2.1 s for 8 singlets, 56 -> 27 GB)
            no, nv = self.mospaces.n_orbs("o1"), self.mospaces.n_orbs("v1")
            n2 = self.mospaces.n_orbs("o2") if self.is_core_valence_separated else no
            big = no * n2 * nv * nv * 8 >= (1 << 30)
            # CVS (packed_cvs.py, virtual pairs only): measured on one B200, c=2 / o=8 / v=106 is still
            # launch-bound (dense 0.09 s, packed 0.15 s for 7 singlets), c=2 / o=22 / v=196 runs 0.29 s
            # dense, 0.19 s packed
            vvvv_bound = (self.block_orders.get("pphh_pphh") == 1
                          and nv >= (160 if self.is_core_valence_separated else 96))
            if not (packable and (big or vvvv_bound)):
                return "dense"
            hf = self.reference_state
            if getattr(hf, "packed_source", None) is None:
                focks = [hf.foo.to_ndarray(), hf.fvv.to_ndarray()]
                if self.is_core_valence_separated:
                    focks.append(hf.fcc.to_ndarray())
                for f in focks:                                          # packed needs canonical orbitals
                    if np.max(np.abs(f - np.diag(np.diag(f)))) > 1e-12:
                        return "dense"
            return "packed"
        return requested

    def _validate_block_orders(self):
        for block, order in self.block_orders.items():
            if order is None:
                continue
            if order < 0:
                raise ValueError("block orders need to be non-negative")
            bra, ket = block.split("_")
            for sp in (bra, ket):
                n_p, n_h = sp.count("p"), sp.count("h")
                if "p" * n_p + "h" * n_h != sp or n_p != n_h:
                    raise ValueError(f"Invalid block {block} for a pp ADC matrix.")
            if bra == ket:
                continue
            inv = self.block_orders.get(f"{ket}_{bra}")
            if inv is None or inv != order:
                raise ValueError(f"{block} and {ket}_{bra} should always have the same order.")
            if self.block_orders.get(f"{bra}_{bra}") is None or self.block_orders.get(f"{ket}_{ket}") is None:
                raise ValueError(f"Can only have a coupling block {block} if both diagonal blocks "
                                 "are in the matrix too.")

    def _init_space_data(self, diagonal):
        self.axis_spaces = {}
        self.axis_lengths = {}
        for block in diagonal.blocks:
            self.axis_spaces[block] = list(diagonal[block].subspaces)
            self.axis_lengths[block] = int(np.prod([self.mospaces.n_orbs(sp)
                                                    for sp in self.axis_spaces[block]]))
        n = sum(self.axis_lengths.values())
        self.shape = (n, n)

    def __repr__(self):
        ret = f"AdcMatrix({self.method.name}, "
        for blk, o in self.block_orders.items():
            ret += f"{blk}={o}, "
        return ret + ")"

    def __len__(self):
        return self.shape[0]

    @property
    def axis_blocks(self):
        return sorted(self.axis_spaces, key=len)

    def diagonal(self):
        return self._diagonal

    @property
    def vector_distribution(self):
        """"single GPU", "replicated" or "slab" (doubles vectors distributed over the GPUs)."""
        eng = self._engine
        if eng is None or eng.world == 1:
            return "single GPU"
        return eng.vectors

    # ------------------------------------------------------------------ interpreter
    def _operand(self, name, ampl):
        if name == "u1":
            return ampl["ph"]
        if name == "u2":
            return ampl["pphh"]
        if name not in self._const:
            if name == "t2":
                t = self.ground_state.t2oo
            elif name in adc_pp.INTERMEDIATES or name in Intermediates.generators:
                t = getattr(self.intermediates, name)
            else:
                t = getattr(self.reference_state, name)
            self._const[name] = t
        return self._const[name]

    def _resolve_constants(self, expr):
        if expr[0] == "einsum":
            for op in expr[2]:
                if isinstance(op, tuple):
                    self._resolve_constants(op)
                elif op not in ("u1", "u2"):
                    self._operand(op, None)
        else:
            self._resolve_constants(expr[1])

    def _eval(self, expr, ampl):
        kind = expr[0]
        if kind == "einsum":
            ops = [self._eval(o, ampl) if isinstance(o, tuple) else self._operand(o, ampl)
                   for o in expr[2]]
            return einsum(expr[1], *ops)
        t = self._eval(expr[1], ampl)
        if kind == "asym":
            for pair in expr[2]:
                t = t.antisymmetrise(*pair)
            return t
        if kind == "sym":
            return t.symmetrise(*expr[2])
        raise ValueError(kind)

    def _terms(self, blk, ampl):
        inb, outb, terms = self._tables[blk]
        if inb is None or inb not in ampl:
            return None, []
        return outb, [(c, self._eval(e, ampl)) for c, e in terms]

    def _make_apply(self, blk):
        def apply(ampl):
            outb, terms = self._terms(blk, ampl)
            if not terms:
                return 0
            res = linear_combination_strict([c for c, _ in terms], [t for _, t in terms])
            return AmplitudeVector(**{outb: res})
        return apply

    def _make_engine_apply(self, blk):
        outb, inb = blk.split("_")

        def apply(ampl):
            if inb not in ampl:
                return 0
            return AmplitudeVector(**{outb: self._engine.block_apply(blk, ampl[inb])})
        return apply

    def block_apply(self, block, tensor):
        if not isinstance(tensor, Tensor):
            raise TypeError("tensor should be an adcc.Tensor")
        with self.timer.record(f"apply/{block}"):
            outblock, inblock = block.split("_")
            ret = self.blocks[block](AmplitudeVector(**{inblock: tensor}))
            return getattr(ret, outblock)

    @contextlib.contextmanager
    def spin_conserving_vectors(self):
        """Promise, for the duration of the block, that every vector handed to ``matvec`` is zero in
        the spin blocks that change M_s - true for the Davidson iterations started from this package's
        singlet / triplet / any guesses (adcc/guess/guess_zero.py:121-161 forbids those blocks, and
        lincomb / preconditioner / matvec keep them zero).  The spin-blocked packed engine then
        multiplies only the matching row class of the vector with each operand block.  No effect on
        other engines; a matvec outside the block makes no assumption."""
        eng = self._engine
        if eng is None or not hasattr(eng, "ms_conserving"):
            yield
            return
        previous = eng.ms_conserving
        eng.ms_conserving = True
        try:
            yield
        finally:
            eng.ms_conserving = previous

    def _matvec_terms(self, v):
        collected = {}
        for blk in self._tables:
            outb, terms = self._terms(blk, v)
            if terms:
                collected.setdefault(outb, []).extend(terms)
        return {outb: linear_combination_strict([c for c, _ in terms], [t for _, t in terms])
                for outb, terms in collected.items()}

    def matvec(self, v):
        """sigma = M v  (adcc/AdcMatrix.py:390-396)."""
        with self.timer.record("matvec"):
            if self.extra_terms:
                return sum(block(v) for block in self.blocks.values())
            if self._engine is not None:
                return self._engine.matvec(v)
            if self._graphs_enabled():
                res = self._graph_matvec([v])
                if res is not None:
                    return res[0]
            return AmplitudeVector(**self._matvec_terms(v))

    # ------------------------------------------------------------------ CUDA-graph replay
    GRAPH_MAX_ELEMENTS = 1 << 25      # dense vectors up to 256 MiB: host launch cost matters
    GRAPH_MAX_CACHED = 8              # distinct (blocks, batch size) signatures kept
    GRAPH_CAPTURE_AFTER = 2           # eager uses of a signature before it is captured

    def _graphs_enabled(self):
        from . import backend as be
        # Opt-in: measured on B200 (tools/diag_small.py) one capture costs about as much as the
        # host time it saves over a single run_adc (9-13 Davidson iterations); it pays when one
        # matrix serves many solves.
        return (os.environ.get("ADCCB_GRAPHS") == "1" and be.graphs_supported()
                and len(self) <= self.GRAPH_MAX_ELEMENTS)

    def _graph_matvec(self, vectors):
        """Replay the captured launch sequence of matvec (one vector) or matvec_batch (several).
        Returns None if no further graph should be cached for this signature."""
        from . import backend as be
        n = len(vectors)
        blocks = tuple(vectors[0].blocks)
        key = (n, blocks)
        graph = self._graphs.get(key)
        lead = () if n == 1 else (n,)
        subspaces = {blk: list(vectors[0][blk].subspaces) for blk in blocks}

        def fill(bufs):
            for blk in blocks:
                for k, v in enumerate(vectors):
                    be.axpby(1.0, v[blk]._contig(), 0.0, bufs[blk] if n == 1 else bufs[blk][k])

        def run(bufs):
            if n == 1:
                ampl = AmplitudeVector(**{blk: Tensor._wrap(self.mospaces, subspaces[blk], bufs[blk])
                                          for blk in blocks})
                return self._matvec_terms(ampl)
            stacked = {blk: Tensor._wrap(self.mospaces, ["B"] + subspaces[blk], bufs[blk])
                       for blk in blocks}
            return self._matvec_stacked(stacked)

        if graph is None:
            # capture costs about ten eager matvecs: only signatures that keep coming back pay off
            seen = self._graph_seen[key] = self._graph_seen.get(key, 0) + 1
            if seen <= self.GRAPH_CAPTURE_AFTER or len(self._graphs) >= self.GRAPH_MAX_CACHED:
                return None
            shapes = {blk: lead + tuple(vectors[0][blk].shape) for blk in blocks}
            graph = self._graphs[key] = be.ReplayGraph(shapes, fill, run)
        outputs = graph.run(fill)
        ret = []
        for k in range(n):
            ret.append(AmplitudeVector(**{
                outb: Tensor._wrap(self.mospaces, res.subspaces[len(lead):],
                                   be.clone(res._data if n == 1 else res._data[k]),
                                   symmetry=res.symmetry if n == 1 else None)
                for outb, res in outputs.items()}))
        return ret

    def rmatvec(self, v):
        return self.matvec(v)

    # ------------------------------------------------------------------ extra terms
    def __iadd__(self, other):
        """In-place addition of an AdcExtraTerm (adcc/AdcMatrix.py:286-311)."""
        if not isinstance(other, AdcExtraTerm):
            return NotImplemented
        if not all(k in self.blocks for k in other.blocks):
            raise ValueError("Can only add to blocks of AdcMatrix that already exist.")
        for sp in other.blocks:
            orig_app, other_app = self.blocks[sp], other.blocks[sp].apply

            def patched_apply(ampl, original=orig_app, other=other_app):
                return sum(app(ampl) for app in (original, other))
            self.blocks[sp] = patched_apply
        other_diagonal = sum(bl.diagonal for bl in other.blocks.values() if bl.diagonal)
        self._diagonal = (self._diagonal + other_diagonal).evaluate()
        self.extra_terms.append(other)
        return self

    def __add__(self, other):
        """matrix + AdcExtraTerm: a copy with the term added (adcc/AdcMatrix.py:313-336)."""
        if not isinstance(other, AdcExtraTerm):
            return NotImplemented
        ret = AdcMatrix(self.method, self.ground_state, block_orders=self.block_orders,
                        intermediates=self.intermediates, diagonal_precomputed=self.diagonal(),
                        doubles_storage=self.doubles_storage)
        ret += other
        return ret

    def __radd__(self, other):
        return self.__add__(other)

    def matvec_host(self, u_host, sigma_host):
        """sigma = M u between PINNED host buffers (dicts {"ph": [no, nv], "pphh": packed
        [npo, npv]} of torch CPU tensors): the end-to-end form of ``matvec`` for callers whose
        vectors live on the host (dense store: tensors of the dense block shapes, plain upload /
        download).  Packed store: the PCIe transfers are pipelined with row slabs of the vvvv GEMM
        (PackedAdcEngine.matvec_host).  Sharded matrices (all ranks must call): either every rank
        passes the SAME buffers in shared pinned host memory (``adcc_b200.shared_host_vector``) and
        moves its own slab over its own PCIe link, or rank 0 passes private buffers and the other
        ranks pass None (upload on rank 0 + NCCL broadcast)."""
        if self._engine is None:
            # dense store: plain upload / matvec / download (the vectors have the dense block shapes)
            from . import backend as be
            with self.timer.record("matvec_host"):
                ampl = AmplitudeVector(**{
                    blk: Tensor._wrap(self.mospaces, self.axis_spaces[blk], be.upload(t)) for blk, t in u_host.items()})
                sigma = self.matvec(ampl)
                for blk, out in sigma_host.items():
                    be.download(sigma[blk]._contig(), out)
            return
        u_host = u_host or {"ph": None, "pphh": None}              # sharded: only rank 0 holds buffers
        sigma_host = sigma_host or {"ph": None, "pphh": None}
        with self.timer.record("matvec_host"):
            self._engine.matvec_host(u_host["ph"], u_host["pphh"], sigma_host["ph"], sigma_host["pphh"])

    # ------------------------------------------------------------------ batched matvec (8(f2))
    def _eval_batched(self, expr, ampl):
        """Like _eval for trial vectors that carry a leading batch axis 'Z' (several vectors stacked):
        returns (tensor, has_batch_axis)."""
        kind = expr[0]
        if kind == "einsum":
            ins, out = expr[1].split("->")
            ins = ins.split(",")
            ops, new_ins, batched = [], [], False
            for idx, o in zip(ins, expr[2]):
                if isinstance(o, tuple):
                    t, b = self._eval_batched(o, ampl)
                else:
                    t, b = self._operand(o, ampl), o in ("u1", "u2")
                ops.append(t)
                new_ins.append(("Z" + idx) if b else idx)
                batched = batched or b
            subs = ",".join(new_ins) + "->" + (("Z" + out) if batched else out)
            return einsum(subs, *ops), batched
        t, batched = self._eval_batched(expr[1], ampl)
        shift = 1 if batched else 0
        if kind == "asym":
            for pair in expr[2]:
                t = t.antisymmetrise(*[p + shift for p in pair])
            return t, batched
        if kind == "sym":
            return t.symmetrise(*[tuple(p + shift for p in pair) for pair in expr[2]]), batched
        raise ValueError(kind)

    def matvec_batch(self, vectors):
        """[M v for v in vectors] with ONE pass over every constant operand: the trial vectors are
        stacked along a leading batch axis, so each term of the block tables is a single GEMM with
        n-times more rows (adcc/AdcMatrix.py:402-408 loops over the list instead)."""
        from . import backend as be
        n = len(vectors)
        if n == 0:
            return []
        blocks = vectors[0].blocks
        if (self._engine is not None and not self.extra_terms and hasattr(self._engine, "matvec_batch")
                and type(self).matvec is AdcMatrix.matvec and all(v.blocks == blocks for v in vectors)):
            with self.timer.record("matvec_batch"):
                return self._engine.matvec_batch(vectors)
        if (n == 1 or self.extra_terms or self._engine is not None
                or os.environ.get("ADCCB_NO_BATCH", "0") == "1"
                or any(v.blocks != blocks for v in vectors)
                or type(self).matvec is not AdcMatrix.matvec):
            return [self.matvec(v) for v in vectors]
        with self.timer.record("matvec_batch"):
            if self._graphs_enabled():
                res = self._graph_matvec(vectors)
                if res is not None:
                    return res
            stacked = {}
            for blk in blocks:
                first = vectors[0][blk]
                data = be.empty((n,) + tuple(first.shape))
                for k, v in enumerate(vectors):
                    be.axpby(1.0, v[blk]._contig(), 0.0, data[k])
                stacked[blk] = Tensor._wrap(self.mospaces, ["B"] + list(first.subspaces), data)
            out = [dict() for _ in range(n)]
            for outb, res in self._matvec_stacked(stacked).items():
                sub = res.subspaces[1:]
                for k in range(n):
                    out[k][outb] = Tensor._wrap(self.mospaces, sub, res._data[k])
            return [AmplitudeVector(**o) for o in out]

    def _matvec_stacked(self, stacked):
        collected = {}
        for blk in self._tables:
            inb, outb, terms = self._tables[blk]
            if inb is None or inb not in stacked:
                continue
            for c, e in terms:
                t, batched = self._eval_batched(e, stacked)
                assert batched
                collected.setdefault(outb, []).append((c, t))
        return {outb: linear_combination_strict([c for c, _ in terms], [t for _, t in terms])
                for outb, terms in collected.items()}

    def __matmul__(self, other):
        if isinstance(other, AmplitudeVector):
            return self.matvec(other)
        if isinstance(other, list) and all(isinstance(e, AmplitudeVector) for e in other):
            return self.matvec_batch(other)
        return NotImplemented

    def block_view(self, block):
        b1, b2 = block.split("_")
        if b1 != b2:
            raise NotImplementedError("Off-diagonal block views not yet implemented.")
        block_orders = {bl: None for bl in self.block_orders}
        block_orders[block] = self.block_orders[block]
        return AdcMatrix(self.method, self.ground_state, block_orders=block_orders,
                         intermediates=self.intermediates, doubles_storage="dense")

    def construct_symmetrisation_for_blocks(self):
        """adcc/AdcMatrix.py:431-472."""
        ret = {}
        if self.is_core_valence_separated:
            ret["pphh"] = lambda v: v.antisymmetrise([(2, 3)])
        else:
            def symmetrise_generic_adc_doubles(invec):
                scratch = invec.antisymmetrise([(0, 1)]).antisymmetrise([(2, 3)])
                return scratch.symmetrise([(0, 1), (2, 3)])
            ret["pphh"] = symmetrise_generic_adc_doubles
        return ret

    def dense_basis(self, axis_blocks=None, ordering="adcc"):
        """Basis in which ``to_ndarray`` expresses the matrix (adcc/AdcMatrix.py:474-558): a list of
        basis vectors, each a list of (index, factor).  Singles: unit vectors; doubles: the
        orthonormal antisymmetrised combinations (+-1/2 on the four images of i>j, a>b; CVS:
        +-1/sqrt(2) on a>b).  ordering: "adcc" (index order), "spin" or "spatial"."""
        if axis_blocks is None:
            axis_blocks = self.axis_blocks
        if not isinstance(axis_blocks, list):
            axis_blocks = [axis_blocks]
        if ordering not in ("adcc", "spin", "spatial"):
            raise ValueError("ordering needs to be 'adcc', 'spin' or 'spatial'")
        if any(b not in ("ph", "pphh") for b in axis_blocks):
            raise NotImplementedError("Blocks other than ph and pphh not implemented")

        def key_of(n_alpha, idx):
            if ordering == "adcc":
                return (idx, idx)
            is_beta = [idx[k] >= n_alpha[k] for k in range(len(idx))]
            spatial = [idx[k] - n_alpha[k] if is_beta[k] else idx[k] for k in range(len(idx))]
            return (is_beta, spatial) if ordering == "spin" else (spatial, is_beta)

        def ordered(vectors, spaces):
            n_alpha = [self.mospaces.n_orbs_alpha(sp) for sp in spaces]
            return sorted(vectors, key=lambda vec: min(key_of(n_alpha, idx) for idx, _ in vec))

        ret = []
        if "ph" in axis_blocks:
            sp = self.axis_spaces["ph"]
            n = [self.mospaces.n_orbs(x) for x in sp]
            ret.extend(ordered([[((i, a), 1)] for i in range(n[0]) for a in range(n[1])], sp))
        if "pphh" in axis_blocks:
            sp = self.axis_spaces["pphh"]
            n = [self.mospaces.n_orbs(x) for x in sp]
            if sp[0] == sp[1] and sp[2] == sp[3]:
                vecs = [[((i, j, a, b), +0.5), ((j, i, a, b), -0.5), ((i, j, b, a), -0.5), ((j, i, b, a), +0.5)]
                        for i in range(n[0]) for j in range(i) for a in range(n[2]) for b in range(a)]
            elif sp[2] == sp[3]:
                f = 1 / np.sqrt(2)
                vecs = [[((i, j, a, b), +f), ((i, j, b, a), -f)]
                        for i in range(n[0]) for j in range(n[1]) for a in range(n[2]) for b in range(a)]
            else:
                vecs = [[((i, j, a, b), 1)] for i in range(n[0]) for j in range(n[1])
                        for a in range(n[2]) for b in range(n[3])]
            ret.extend(ordered(vecs, sp))
        return ret

    def to_ndarray(self, out=None):
        """The matrix as a dense array in the ``dense_basis`` (adcc/AdcMatrix.py:560-644; debugging /
        visualisation: one matvec per basis vector).  No spin symmetry is imposed: the spectrum
        contains singlets and triplets and, for restricted references, spin-mixed combinations."""
        if any(b not in ("ph", "pphh") for b in self.axis_blocks):
            raise NotImplementedError("Blocks other than ph and pphh not implemented")
        if "ph" not in self.axis_blocks:
            raise NotImplementedError("Block 'ph' needs to be present")
        basis = {b: self.dense_basis(b) for b in self.axis_blocks}
        mat_len = sum(len(v) for v in basis.values())
        if out is None:
            out = np.zeros((mat_len, mat_len))
        elif out.shape != (mat_len, mat_len):
            raise ValueError(f"Output array has shape ({out.shape[0]}, {out.shape[1]}), but shape "
                             f"({mat_len}, {mat_len}) is required.")
        else:
            out[:] = 0
        shapes = {b: tuple(self.mospaces.n_orbs(sp) for sp in self.axis_spaces[b]) for b in self.axis_blocks}
        # basis vectors as columns of a (dense length) x (basis length) array per block
        expand = {}
        for b in self.axis_blocks:
            E = np.zeros((int(np.prod(shapes[b])), len(basis[b])))
            for j, vec in enumerate(basis[b]):
                for idx, val in vec:
                    E[np.ravel_multi_index(idx, shapes[b]), j] = val
            expand[b] = E
        offs, off = {}, 0
        for b in self.axis_blocks:
            offs[b] = off
            off += len(basis[b])
        chunk = 16
        for b in self.axis_blocks:
            for j0 in range(0, len(basis[b]), chunk):
                cols = range(j0, min(j0 + chunk, len(basis[b])))
                vectors = []
                for j in cols:
                    data = {}
                    for blk in self.axis_blocks:
                        arr = expand[b][:, j].reshape(shapes[b]) if blk == b else np.zeros(shapes[blk])
                        data[blk] = Tensor.from_ndarray(self.mospaces, self.axis_spaces[blk], arr)
                    vectors.append(AmplitudeVector(**data))
                for j, res in zip(cols, self @ vectors):
                    for blk in res.blocks:
                        dense = res[blk].to_ndarray().reshape(-1)
                        out[offs[blk]:offs[blk] + len(basis[blk]), offs[b] + j] = expand[blk].T @ dense
        return out


class AdcMatrixShifted(AdcMatrix):
    """``matrix @ v + shift * v`` (adcc/AdcMatrix.py:649-693)."""

    def __init__(self, matrix, shift=0.0):
        super().__init__(matrix.method, matrix.ground_state, block_orders=matrix.block_orders,
                         intermediates=matrix.intermediates, doubles_storage=matrix.doubles_storage)
        self.shift = shift

    def matvec(self, in_ampl):
        return super().matvec(in_ampl) + self.shift * in_ampl

    def block_apply(self, block, in_vec):
        ret = super().block_apply(block, in_vec)
        inblock, outblock = block.split("_")
        if inblock == outblock:
            ret = ret + self.shift * in_vec
        return ret

    def diagonal(self):
        return super().diagonal() + self.shift

    def to_ndarray(self, out=None):
        dense = super().to_ndarray(out)
        dense += self.shift * np.eye(dense.shape[0])
        return dense

    def block_view(self, block):
        raise NotImplementedError("Block-view not yet implemented for shifted ADC matrices.")


class AdcMatrixProjected(AdcMatrix):
    """``P @ M @ P`` with P a projector onto selected excitation blocks, e.g. ["cv", "ccvv", "ocvv"]
    (adcc/AdcMatrix.py:696-780)."""

    def __init__(self, matrix, excitation_blocks, core_orbitals=None, outer_virtuals=None):
        from .projection import Projector, SubspacePartitioning
        for sp in excitation_blocks:
            if not any(len(sp) == len(ax_sp) for ax_sp in matrix.axis_spaces.values()):
                raise ValueError(f"Invalid partition block {sp}.")
        super().__init__(matrix.method, matrix.ground_state, block_orders=matrix.block_orders,
                         intermediates=matrix.intermediates, doubles_storage="dense")
        partitioning = SubspacePartitioning(matrix.mospaces, core_orbitals, outer_virtuals)
        self.projectors = {}
        for block in matrix.axis_spaces:
            keep = [sp for sp in excitation_blocks if len(sp) == len(matrix.axis_spaces[block])]
            self.projectors[block] = Projector(matrix.axis_spaces[block], partitioning, keep)

    def apply_projection(self, in_ampl):
        return AmplitudeVector(**{block: self.projectors[block] @ in_ampl[block] for block in in_ampl.keys()})

    def matvec(self, in_ampl):
        return self.apply_projection(super().matvec(self.apply_projection(in_ampl)))

    def block_apply(self, block, in_vec):
        outblock, inblock = block.split("_")
        ret = super().block_apply(block, self.projectors[inblock].apply(in_vec))
        return self.projectors[outblock].apply(ret)

    def diagonal(self):
        blocks = {}
        for block, diagblock in super().diagonal().items():
            P = self.projectors[block]
            one = diagblock.ones_like()
            blocks[block] = P @ diagblock + 100000 * (one - P @ one)
        return AmplitudeVector(**blocks)

    def block_view(self, block):
        raise NotImplementedError("Block-view not yet implemented for projected ADC matrices.")
