"""ADC(3) operands of the packed engine, built without ever forming a dense ovvv or vvvv block.

What the third-order matrix needs beyond ADC(2)-x (adcc/adc_pp/matrix.py:490-532, 606-618):

  * ``adc3_m11`` (matrix.py:992-1036), the (ov x ov) singles block with the intermediates
    ``adc3_i1`` / ``adc3_i2`` (:945-989), ``td2`` (adcc/LazyMp.py:136-151), ``mp2_diffdm`` incl.
    ``ts2_ov`` (adcc/GroundState.py:285-324, adcc/LazyMp.py:72-104), ``t2eri`` pi3 / pi4 / pi5
    (GroundState.py:219-241) and ``t2sq`` (:1548-1550);
  * the coupling intermediates ``adc3_pia`` (:1070-1084, pi1 / pi2) and ``adc3_pib`` (:1087-1094,
    pi6 / pi7) in the GEMM layouts of the packed engine;
  * operands of the four three-operand second-order coupling terms (:498-499, 525, 529).

At nocc=40 / nvirt=400 the dense vvvv block is 204.8 GB and ovvv 20.5 GB, so every term that
touches them runs on the pair-packed operands W[AB,CD], V2[a,(j,BC)] the engine holds anyway:

  pi5, pi3      packed GEMMs (the vvvv / oooo terms of the matvec applied to t2), unpacked to o^2 v^2
  pi4           the ovov-term GEMM of the matvec applied to t2
  ts2, pi1, pi6 GEMMs over the packed pair index of V2 / T2
  <ia||bc> p[ic], <ib||ac> x[jc]     antisymmetric pair matvec (adccb_pair_matvec) on V2
  pi7 (o^2 v^4) virtual-index slabs: exchange-ordered ovvv slab (adccb_ovvv_exchange_slab) x X(t2),
                folded straight into the packed pib operand
  <ac||bd> p[cd], -t2sq[idjc] <ac||bd>, 1/2 <ic||jd> t2[klac] t2[klbd]   (o^2 v^4 each)
                virtual-index slabs: exchange-ordered vvvv slab (adccb_vvvv_exchange_slab) and the
                matching slab of t2.t2 built by a GEMM, both consumed by DMMA GEMMs

Everything of size o^2 v^2 or smaller goes through the generic Tensor / einsum layer (the equations
of adcc_b200/adc_pp.py), with the pieces above injected through ``hf.big_eri_ops`` and the caches of
LazyMp.  Works for any operand source (DenseSource: real molecules; SyntheticSource: BASELINE
configs[4]), which is how it is tested against the dense builders and the oracle.
"""
from . import backend as be
from . import block as b
from .functions import einsum
from .Tensor import Tensor

O, V = "o1", "v1"


def _slab(n_total, bytes_per_index, budget=8 << 30):
    """Number of virtual indices per slab so that one slab temporary stays below ``budget`` bytes."""
    return max(1, min(n_total, int(budget // max(bytes_per_index, 1))))


class PackedEriOps:
    """hf.big_eri_ops on pair-packed operands (see adc_pp.DenseEriOps for the dense statement)."""

    def __init__(self, mospaces, no, nv, src, W, V2e, T2p, j0=0, nj=None, shard=(0, 1), group=None):
        """``V2e``: [a,(jl,BC)] for the occupied slab j0 .. j0+nj-1 this rank owns; ``W``: whatever
        ``src.vvvv_exchange_slab`` needs (the complete pair-packed block, or None for a generating
        source).  With shard = (rank, world) the virtual-index slab loops are dealt out round-robin
        and every result is completed by an all-reduce over NVLink."""
        self.ms, self.no, self.nv = mospaces, no, nv
        self.npo, self.npv = be.n_pairs(no), be.n_pairs(nv)
        self.src, self.W, self.V2e, self.T2p = src, W, V2e, T2p
        self.j0, self.nj = j0, (no if nj is None else nj)
        self.rank, self.world = shard
        self.group = group
        self.flops = 0.0

    def _t(self, data, spaces):
        return Tensor._wrap(self.ms, spaces, data)

    def _sum(self, data):
        if self.world > 1:
            be.comm_all_reduce(data, group=self.group)
        return data

    def _my_slabs(self, n_total, na):
        return [a0 for k, a0 in enumerate(range(0, n_total, na)) if k % self.world == self.rank]

    # <ia||bc> p[ic] -> [a,b]        (matrix.py:951)
    def ovvv_iabc_ic(self, p_ov):
        nv, nj = self.nv, self.nj
        out = be.zeros((nv, nv))
        if nj > 0:
            be.pair_matvec(self.V2e, nv, n_out=nv, n_red=nj, row_stride_out=nj, row_stride_red=1,
                           X=p_ov._contig()[self.j0:self.j0 + nj], x_by_out=False, out=out)
        return self._t(self._sum(out), [V, V])

    # <ib||ac> x[jc] -> [i,a,j,b]    (matrix.py:1014)
    def ovvv_ibac_jc(self, x_ov):
        no, nv, nj, j0 = self.no, self.nv, self.nj, self.j0
        x = x_ov._contig()
        res = be.zeros((no, nv, no, nv))                       # [j, b, i, a]; this rank fills i in its slab
        if nj > 0:
            tmp = be.empty((nv * nj, nv))
            for j in range(no):
                xj = x[j].reshape(1, nv).expand(nv * nj, nv)   # the same vector for every row (b, il)
                be.pair_matvec(self.V2e, nv, n_out=nv * nj, n_red=1, row_stride_out=1, row_stride_red=0,
                               X=xj, x_by_out=True, out=tmp)
                be.copy2d(res[j].view(nv, no * nv)[:, j0 * nv:(j0 + nj) * nv], tmp.view(nv, nj * nv), nv, nj * nv)
        return self._t(self._sum(res), [O, V, O, V]).transpose((2, 3, 0, 1))

    # <ac||bd> p[cd] -> [a,b]        (matrix.py:953)
    def vvvv_acbd_cd(self, p_vv):
        nv = self.nv
        p = p_vv._contig().reshape(1, nv * nv)
        out = be.zeros((nv, nv))
        na = _slab(nv, 8 * nv ** 3)
        for a0 in self._my_slabs(nv, na):
            n = min(na, nv - a0)
            Z = self.src.vvvv_exchange_slab(self.W, a0, n).view(n * nv, nv * nv)
            be.gemm(Z, p, out=out[a0:a0 + n].view(n * nv, 1))
            self.flops += 2.0 * n * nv * nv * nv
        return self._t(self._sum(out), [V, V])

    # 1/2 <ic||jd> (t2[klac] t2[klbd]) - t2sq[idjc] <ac||bd> -> [i,a,j,b]   (matrix.py:1029-1033)
    def m11_v4(self, ovov, t2, t2sq):
        no, nv, npo = self.no, self.nv, self.npo
        B1 = be.contiguous(t2sq._data.permute(0, 2, 3, 1)).view(no * no, nv * nv)     # t2sq[i,d,j,c] -> [(i,j),(c,d)]
        B2 = be.contiguous(ovov._data.permute(0, 2, 1, 3)).view(no * no, nv * nv)     # <ic||jd>     -> [(i,j),(c,d)]
        # t2 as [(a,c), KL] over canonical occupied pairs: t2[klac] t2[klbd] = 2 sum_{K<L}
        Uh = be.packed_expand(2, self.T2p, no, nv).view(nv, npo, nv)                    # [a, KL, c]
        T2ac = be.contiguous(Uh.permute(0, 2, 1)).view(nv * nv, npo)
        del Uh
        res = be.zeros((no * no, nv * nv))                                              # [(i,j),(a,b)]
        na = _slab(nv, 3 * 8 * nv ** 3, budget=24 << 30)
        if self.world > 1:
            na = max(1, min(na, -(-nv // (4 * self.world))))      # enough slabs to balance the ranks
        for a0 in self._my_slabs(nv, na):
            n = min(na, nv - a0)
            cols = res[:, a0 * nv:(a0 + n) * nv]
            Z = self.src.vvvv_exchange_slab(self.W, a0, n).view(n * nv, nv * nv)       # [(a,b),(c,d)] = <ac||bd>
            be.gemm(B1, Z, out=cols, alpha=-1.0)
            del Z
            Q = be.gemm(T2ac[a0 * nv:(a0 + n) * nv], T2ac, alpha=2.0)                  # [(a,c),(b,d)]
            Qx = be.permute(Q.view(n, nv, nv, nv), (0, 2, 1, 3)).view(n * nv, nv * nv)  # [(a,b),(c,d)]
            del Q
            be.gemm(B2, Qx, out=cols, alpha=0.5, beta=1.0)
            del Qx
            self.flops += 2 * 2.0 * no * no * n * nv ** 3 + 2.0 * n * nv * nv * nv * npo
        return self._t(self._sum(res).view(no, no, nv, nv), [O, O, V, V]).transpose((0, 2, 1, 3))


def build_adc3_operands(hf, mp, im, src, W, Op, Y, j0=0, nj=None, shard=(0, 1), layout=None, W_slab=None):
    """Returns the dict of operands the packed engine needs at third order.

    ``Op``, ``Y``: pair-packed oooo and the complete [(j,b),(k,c)] ovov operand.  ``W``: the complete
    pair-packed vvvv block (single GPU / dense sources) or None when ``src`` generates exchange-ordered
    slabs itself.  ``j0``, ``nj``: occupied slab of the ovvv-type operands this rank owns.
    Sharded build (shard = (rank, world), ``layout`` = SlabLayout, ``W_slab`` = this rank's AB rows of
    W): pi5 is assembled from the ranks' column slabs, the ovvv contractions run on the rank's
    occupied slab, the o^2 v^4 slab loops are dealt out over the ranks; partial results are summed by
    all-reduces, the pib operands are produced for the rank's own slab only (no communication)."""
    ms = hf.mospaces
    no, nv = ms.n_orbs(O), ms.n_orbs(V)
    npo, npv = be.n_pairs(no), be.n_pairs(nv)
    nj = no if nj is None else nj
    rank, world = shard
    group = layout.group if layout is not None else None
    flops = 0.0

    def T(data, spaces):
        return Tensor._wrap(ms, spaces, data)

    def seed(cache, key, tensor):
        tensor._contig()
        tensor.set_immutable()
        cache[key] = tensor
        return tensor

    def allsum(data):
        if world > 1:
            be.comm_all_reduce(data, group=group)
        return data

    V2e = src.ovvv_ajbc(j0, nj)               # [a,(jl,BC)] = <ja||bc>, occupied slab of this rank
    G1e = src.ooov_ijkb()                     # [i,(JK,b)] = <jk||ib>
    G2e = src.ooov_ijak()                     # [(IJ,a),k] = <ij||ka>
    V2x = V2e if world == 1 else None         # operand of src.ovvv_exchange_slab (complete, or generated)
    if world > 1 and not getattr(src, "generates_exchange_slabs", False):
        V2x = src.ovvv_ajbc()
    t2 = mp.t2oo
    T2p = be.packed_pack(t2._contig(), no, nv)

    # ---- t2eri pi5 / pi3 / pi4 (GroundState.py:219-241) into LazyMp's cache
    if world == 1:
        p5 = be.gemm(T2p, W, alpha=2.0)
    else:
        slab = be.zeros((npo, layout.wmax))
        if layout.w > 0:
            be.gemm(T2p, W_slab, out=slab[:, :layout.w], alpha=2.0)
        p5 = layout.gather(slab)
        del slab
    seed(mp._cache, ("t2eri", b.oovv + b.vv), T(be.packed_unpack(p5, no, nv), [O, O, V, V]))
    del p5
    p3 = be.gemm(Op, be.permute(T2p, (1, 0)), alpha=2.0)
    seed(mp._cache, ("t2eri", b.oovv + b.oo), T(be.packed_unpack(p3, no, nv), [O, O, V, V]))
    del p3
    X2 = be.packed_expand(0, T2p, no, nv)                                   # [(i,a),(k,c)] = t2[ikac]
    T4 = be.gemm(X2, Y)                                                     # [(j,a),(i,b)] = t2[jkac] <kb||ic>
    seed(mp._cache, ("t2eri", b.oovv + b.ov),
         T(be.contiguous(T4.view(no, nv, no, nv).permute(2, 0, 1, 3)), [O, O, V, V]))
    del T4
    flops += 2.0 * npo * npv * npv / world + 2.0 * npo * npo * npv + 2.0 * (no * nv) ** 3
    # ---- ts2_ov = 1/2 ph_pphh_1(t2) / (-df)   (LazyMp.py:72-104)
    Ue = be.packed_expand(1, T2p, no, nv)                                   # [i,(j,BC)]
    s = be.zeros((no, nv))
    if nj > 0:
        be.gemm(Ue[:, j0 * npv:(j0 + nj) * npv], V2e, out=s, alpha=2.0)     # t2[ijbc] <ja||bc>, j in slab
    allsum(s)
    Uh = be.packed_expand(2, T2p, no, nv)                                   # [a,(JK,b)]
    be.gemm(G1e, Uh, out=s, alpha=2.0, beta=1.0)                            # <jk||ib> t2[jkab]
    seed(mp._cache, ("ts2_ov", False), (0.5 * T(s, [O, V])) / (-1.0 * mp.df(b.ov)))

    # ---- singles block: the generic equations with the ovvv / vvvv contractions injected
    ops = PackedEriOps(ms, no, nv, src, W, V2e, T2p, j0, nj, shard, group)
    hf.big_eri_ops = ops
    try:
        M1 = im.adc3_m11._contig().reshape(no * nv, no * nv)
    finally:
        hf.big_eri_ops = None
    flops += ops.flops

    # ---- adc3_pia in pair-packed form [IJ,a,k]  (matrix.py:1070-1084)
    pi2 = mp.t2eri(b.ooov, b.ov)                                            # t2[ilab] <lk||jb> -> [i,j,k,a]
    base = hf.ooov - 2.0 * pi2.antisymmetrise(0, 1)
    piaP = be.permute(be.pair_select(base._contig(), 0), (0, 2, 1))         # [IJ,a,k]
    pi1 = be.zeros((npo, nv, no))                                           # [IJ,a,k] = t2[ijbc] <ka||bc>
    if nj > 0:
        part = be.gemm(T2p, V2e.view(nv * nj, npv), alpha=2.0)              # [IJ,(a,kl)], k in slab
        be.copy2d(pi1.view(npo * nv, no)[:, j0:j0 + nj], part.view(npo * nv, nj), npo * nv, nj)
        del part
    allsum(pi1)
    be.axpby(-0.5, pi1, 1.0, piaP)
    del pi1, base
    G2 = piaP.view(npo * nv, no)                                            # [(IJ,a),k]
    G1 = be.permute(piaP.view(npo, nv, no), (2, 0, 1)).reshape(no, npo * nv)   # [i,(JK,b)]
    flops += 2.0 * npo * npv * nj * nv

    # ---- adc3_pib of the rank's occupied slab in the V2 layout [a,(il,BC)]  (matrix.py:1087-1094)
    pibV2 = be.clone(V2e)
    Gt2 = be.contiguous(be.permute(G2e.view(npo, nv, no), (1, 2, 0))[:, j0:j0 + nj]).reshape(nv * nj, npo)
    S2 = pibV2.view(nv * nj, npv)                                           # rows (a, il)
    if nj > 0:
        be.gemm(Gt2, be.permute(T2p, (1, 0)), out=S2, alpha=-1.0, beta=1.0)  # -1/2 pi6, [(a,il),JK] = <jk||ia>
    del Gt2
    X2s = X2[j0 * nv:(j0 + nj) * nv]                                        # rows (il, b) of X(t2)
    na = _slab(nv, 2 * 8 * nv * no * nv, budget=4 << 30)
    for a0 in (range(0, nv, na) if nj > 0 else ()):                         # + 2 A_bc pi7, pi7 = t2[ijbd] <jc||ad>
        n = min(na, nv - a0)
        Vx = src.ovvv_exchange_slab(V2x, a0, n).view(n * nv, no * nv)      # [(a,c),(j,d)]
        P7 = be.gemm(Vx, X2s)                                               # [(a,c),(il,b)] = pi7[i,a,b,c]
        del Vx
        off = be.int64_tensor([al * nv * nj * nv + il * nv for al in range(n) for il in range(nj)])
        rows = be.int64_tensor([(a0 + al) * nj + il for al in range(n) for il in range(nj)])
        be.packed_fold_virt(S2, nv, nj * nv, P7, off, alpha=-1.0, rows=rows)
        del P7
    flops += 2.0 * nv * nv * (no * nv) * (nj * nv) + 2.0 * nv * nj * npo * npv
    del X2, X2s, V2x
    V2 = pibV2.view(nv, nj * npv)                                           # [a,(jl,BC)]
    del S2

    # ---- operands of the three-operand second-order coupling terms (matrix.py:498-499, 525, 529)
    ooov = hf.ooov._contig()                                                # [i,j,k,a]
    extras = {
        "V2e": V2e,                                                                  # [a,(jl,BC)], slab
        "t2e": Ue,                                                                   # [j,(l,BC)] = t2[jl,BC]
        "t2eM": be.permute(Ue.view(no, no, npv)[j0:j0 + nj], (0, 2, 1)).reshape(nj * npv, no),   # [(jl,AB),l]
        "T2uh": Uh,                                                                  # [b,(JK,d)] = t2[JK,b,d]
        "T2h": be.permute(Uh.view(nv, npo, nv), (1, 0, 2)).reshape(npo * nv, nv),    # [(IJ,a),c]
        "G3": be.permute(ooov, (0, 3, 1, 2)).reshape(no * nv, no * no),              # [(i,a),(j,k)] = <ij||ka>
        "G4": be.permute(ooov, (0, 2, 1, 3)).reshape(no * no, no * nv),              # [(l,i),(k,c)] = <lk||ic>
    }
    if no * no * nv * nv * 8 >= (1 << 30):
        # large problems: the o^2 v^2 intermediates of the build (pi3 / pi4 / pi5, td2, t2sq, cached GEMM
        # layouts of the ERI blocks: ~2 GB each at nocc=40 / nvirt=400) are not needed by the matvec;
        # give the memory back to the Davidson subspace.  They are rebuilt on demand.
        mp._cache.clear()
        im.cached_tensors.clear()
        for t in list(getattr(hf, "_eri", {}).values()) + list(getattr(hf, "_fock", {}).values()):
            t._layouts.clear()
        be.release_cached_memory()
    return {"M1": M1, "G1": G1, "G2": G2, "V2": V2, "extras": extras, "build_flops": flops}
