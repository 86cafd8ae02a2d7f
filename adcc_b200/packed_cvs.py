"""Virtual-pair-packed doubles store and matvec for the core-valence separated variants
CVS-ADC(2) and CVS-ADC(2)-x.

The CVS doubles amplitudes u[i,J,a,b] (i valence occupied "o1", J core occupied "o2") are
antisymmetric in (a,b) only (adcc/AdcMatrix.py:431-436, adcc/guess/guess_zero.py:143-171), and
<ab||cd> is antisymmetric in both pairs.  Keeping a<b (and c<d) turns

    1/2 u[iJcd] <ab||cd>                                     (adcc/adc_pp/matrix.py:262)

into ONE GEMM  S[(i,J),AB] = sum_CD U[(i,J),CD] W[AB,CD]  on the same packed W[AB,CD] the generic
engine uses (4x fewer flops, 4x less memory than the dense block), halves every Davidson vector
and makes the explicit doubles antisymmetrisation the identity.

Blocks evaluated here (everything through the C ABI: adccb_dgemm incl. its N-major B form,
adccb_strided_gather, adccb_pair_select, adccb_block_scatter, adccb_packed_fold_virt,
adccb_elementwise):
    adcc/adc_pp/matrix.py:136-143   pphh_pphh_0 (cvs)
    adcc/adc_pp/matrix.py:249-264   pphh_pphh_1 (cvs)
    adcc/adc_pp/matrix.py:279-285   ph_pphh_1   (cvs)
    adcc/adc_pp/matrix.py:297-304   pphh_ph_1   (cvs)
    adcc/adc_pp/matrix.py:378-391   ph_ph_2     (cvs; the matrix's table interpreter: three GEMMs
                                                 on nc x nv singles)
Requires canonical orbitals (diagonal Fock blocks), as the generic packed engine does.
"""
import numpy as np

from . import adc_pp
from . import backend as be
from .AmplitudeVector import AmplitudeVector
from .packed import PackedDoubles
from .Tensor import Tensor

SQ2 = float(np.sqrt(2.0))


def _pair_columns(nv):
    """Flat positions a*nv+b and b*nv+a of every pair a<b in pair order AB = b(b-1)/2 + a."""
    b, a = np.tril_indices(nv, -1)                # rows b > cols a, ordered by b then a: pair order
    return a * nv + b, b * nv + a


class CvsPackedDoubles(PackedDoubles):
    """pphh tensor of the CVS variants, antisymmetric in (2,3), stored as U[(i,J), pair(a<b)]."""

    #: every stored element stands for 2 elements of the dense tensor in a Frobenius product
    dot_weight = 2.0
    _storage = "packed_cvs"
    _cols = {}

    @classmethod
    def zeros(cls, mospaces, subspaces, symmetry=None):
        if subspaces[2] != subspaces[3]:
            raise ValueError("CvsPackedDoubles needs equal virtual subspaces")
        no, nc, nv = (mospaces.n_orbs(subspaces[k]) for k in (0, 1, 2))
        return cls._wrap(mospaces, subspaces, be.zeros((no * nc, be.n_pairs(nv))), symmetry)

    @classmethod
    def from_dense(cls, tensor):
        """0.5 (T - T^(ab)) packed: antisymmetrise (2,3) + keep a<b."""
        no, nc, nv = tensor.shape[0], tensor.shape[1], tensor.shape[2]
        d = tensor._contig()
        asym = be.permute(d, (0, 1, 3, 2), alpha=-0.5, add=d, gamma=0.5)
        return cls._wrap(tensor.mospaces, tensor.subspaces,
                         be.pair_select(asym, 2).reshape(no * nc, be.n_pairs(nv)), tensor.symmetry)

    @property
    def _nc(self): return self.mospaces.n_orbs(self.subspaces[1])
    @property
    def shape(self): return (self._no, self._nc, self._nv, self._nv)
    @property
    def size(self): return self._no * self._nc * self._nv ** 2

    def __repr__(self):
        return f"CvsPackedDoubles({self.space}, packed={self.packed_shape})"

    @classmethod
    def _columns(cls, nv):
        key = (nv, str(be.device()))
        if key not in cls._cols:
            ab, ba = _pair_columns(nv)
            cls._cols[key] = (be.int64_tensor(ab), be.int64_tensor(ba))
        return cls._cols[key]

    def to_dense(self):
        no, nc, nv = self._no, self._nc, self._nv
        ab, ba = self._columns(nv)
        dense = be.zeros((no * nc, nv * nv))
        be.block_scatter_add(dense, self._contig(), None, ab, assign=True)
        be.block_scatter_add(dense, self._contig(), None, ba, alpha=1.0 if self.parity > 0 else -1.0, assign=True)
        return Tensor._wrap(self.mospaces, self.subspaces, dense.reshape(no, nc, nv, nv), self.symmetry)

    def set_from_ndarray(self, array, symmetry_tolerance=None):
        self._require_mutable()
        array = np.asarray(array, dtype=float)
        if array.shape != self.shape:
            raise ValueError(f"Shape mismatch: tensor {self.shape} vs array {array.shape}")
        tol = 1e-12 if symmetry_tolerance is None else symmetry_tolerance
        if np.max(np.abs(array + array.transpose(0, 1, 3, 2))) > tol:
            raise ValueError("array is not antisymmetric in (2,3)")
        self._data = type(self).from_dense(
            Tensor._wrap(self.mospaces, self.subspaces, be.from_numpy(array)))._data
        return self

    def dot(self, other):
        if isinstance(other, Tensor):
            self._check_same(other)
            return self.dot_weight * be.dot(self._contig(), other._contig())
        for o in other:
            self._check_same(o)
        return be.dot_many(self._contig(), [o._contig() for o in other], scale=self.dot_weight)

    def _locate(self, index):
        i, j, a, b = self._norm_index(index)
        if a == b:
            return None, 0.0
        sign = 1.0 if a < b else -1.0
        a, b = min(a, b), max(a, b)
        return (i * self._nc + j, b * (b - 1) // 2 + a), sign

    def _symm(self, perms, sign):
        from .Tensor import _parse_perm_args
        perms = sorted(tuple(sorted(p)) for p in _parse_perm_args(perms))
        if sign < 0 and perms == [(2, 3)]:
            return self            # 1/2 (T - P T) = T for the antisymmetric pair
        return self.to_dense()._symm(perms, sign)


class PackedCvsEngine:
    """CVS-ADC(2) / CVS-ADC(2)-x matvec on the virtual-pair-packed doubles store (module docstring)."""

    world, rank = 1, 0
    vectors = "single GPU"
    diagonal_ph = None
    _extra3 = None

    def __init__(self, matrix, shard=None, vectors=None):
        if shard is not None and shard[1] > 1:
            raise NotImplementedError("the CVS packed engine runs on one GPU")
        hf = matrix.reference_state
        self.matrix = matrix
        self.mospaces = hf.mospaces
        self.no, self.nc, self.nv = (self.mospaces.n_orbs(s) for s in ("o1", "o2", "v1"))
        no, nc, nv = self.no, self.nc, self.nv
        self.npv = be.n_pairs(nv)
        self.R = no * nc
        self.order_dd = matrix.block_orders["pphh_pphh"]
        for blk in ("foo", "fcc", "fvv"):
            f = getattr(hf, blk).to_ndarray()
            if np.max(np.abs(f - np.diag(np.diag(f)))) > 1e-10:
                raise NotImplementedError("the packed CVS engine needs canonical orbitals")
        self._ph_ph = matrix._make_apply("ph_ph")            # nc x nv singles: table interpreter
        for _, expr in matrix._tables["ph_ph"][2]:
            matrix._resolve_constants(expr)

        # ---- couplings (matrix.py:279-285, 297-304)
        ovvv = be.pair_select(hf.ovvv._contig(), 2)                                     # [j,a,BC]
        self.V2 = be.permute(ovvv, (1, 0, 2)).reshape(nv, no * self.npv)                # [a,(j,BC)] = <ja||bc>
        del ovvv
        occv = hf.occv._contig()                                                        # [j,K,I,b]
        self.Gc = be.permute(occv, (2, 0, 1, 3)).reshape(nc, self.R * nv)               # [I,(jK,b)]
        self.Gb = be.permute(occv, (0, 1, 3, 2)).reshape(self.R * nv, nc)               # [(jI,b),K]

        # ---- doubles-doubles (matrix.py:136-143, 249-264)
        def packed_diagonal(order):
            d = adc_pp.diagonal("cvs", "pphh_pphh", order, hf, matrix.ground_state, matrix.intermediates)
            return be.pair_select(d.pphh.evaluate()._contig(), 2).reshape(self.R, self.npv)
        self.D0 = packed_diagonal(0)                         # e_a + e_b - e_i - e_J: the Fock terms
        self.diagonal_pphh = CvsPackedDoubles._wrap(
            self.mospaces, ["o1", "o2", "v1", "v1"], self.D0 if self.order_dd == 0 else packed_diagonal(1))
        self.diagonal_pphh.parity = +1
        self.flops = {"ph_pphh": 2.0 * nc * nv * (self.R * nv + no * self.npv),
                      "pphh_ph": 2.0 * nc * nv * (self.R * nv + no * self.npv)}
        if self.order_dd == 1:
            self.W = be.even_pitch(be.packed_pack(hf.vvvv._contig(), nv, nv))           # [AB,CD], TMA-ready rows
            self.O = be.contiguous(hf.ococ._contig().reshape(self.R, self.R))            # [(iJ),(lK)]
            self.Ycv = be.permute(hf.cvcv._contig(), (2, 1, 0, 3)).reshape(nc * nv, nc * nv)   # [(J,b),(K,c)]
            self.Yov = be.permute(hf.ovov._contig(), (0, 3, 2, 1)).reshape(no * nv, no * nv)   # [(i,a),(k,c)]
            self._off_T = be.int64_tensor([i * nv * nc * nv + J * nv for i in range(no) for J in range(nc)])
            self.flops["pphh_pphh/vvvv"] = 2.0 * self.R * self.npv ** 2
            self.flops["pphh_pphh/ococ"] = 2.0 * self.R ** 2 * self.npv
            self.flops["pphh_pphh/cvcv+ovov"] = 2.0 * (no * nv) * (nc * nv) * (nc * nv + no * nv)
        self._off_rows = be.int64_tensor([r * nv * nv for r in range(self.R)])
        self.flops_per_matvec = float(sum(self.flops.values()))
        self.build_flops = 0.0

    # ------------------------------------------------------------------ helpers
    def _wrap_ph(self, data):
        return Tensor._wrap(self.mospaces, ["o2", "v1"], data)

    def _wrap_pphh(self, data):
        return CvsPackedDoubles._wrap(self.mospaces, ["o1", "o2", "v1", "v1"], data)

    def _dense(self, U):
        """[(i,J),a,b] antisymmetric expansion of the packed rows."""
        return self._wrap_pphh(U).to_dense()._data.reshape(self.R, self.nv, self.nv)

    def _packed(self, t):
        t = t if isinstance(t, CvsPackedDoubles) else CvsPackedDoubles.from_dense(t)
        return t._contig()

    # ------------------------------------------------------------------ block applies
    def apply_ph_ph(self, u1):
        return self._ph_ph(AmplitudeVector(ph=u1)).ph._contig()

    def apply_ph_pphh(self, U, Ud=None):
        """+ sqrt2 <jK||Ib> u[jKab] - 1/sqrt2 u[jIbc] <ja||bc>   (matrix.py:279-285)."""
        no, nc, nv = self.no, self.nc, self.nv
        Ud = self._dense(U) if Ud is None else Ud
        # sum_{(jK),b} G[I,(jK,b)] u[jK,a,b] = - sum G[I,(jK,b)] Ud[(jK,b),a]: N-major B, no transpose
        s1 = be.gemm(self.Gc, Ud.view(self.R * nv, nv).t(), alpha=-SQ2)
        Ue = be.permute(U.view(no, nc, self.npv), (1, 0, 2)).reshape(nc, no * self.npv)   # [I,(j,BC)]
        be.gemm(Ue, self.V2, out=s1, alpha=-SQ2, beta=1.0)        # -1/sqrt2 * 2 sum_{b<c}
        return s1

    def apply_pphh_ph(self, u1, S):
        """+ sqrt2 A_ab(<jI||Kb> u[Ka]) - 1/sqrt2 u[Ic] <jc||ab>   (matrix.py:297-304)."""
        no, nc, nv = self.no, self.nc, self.nv
        x = u1._contig()                                                                   # [K,a]
        c1 = be.gemm(x, self.V2.t())                                                       # [I,(j,AB)]
        be.permute(c1.view(nc, no, self.npv), (1, 0, 2), alpha=-1.0 / SQ2, out=S.view(no, nc, self.npv), beta=1.0)
        c2 = be.gemm(self.Gb, x.t())                                                       # [(jI,b),a] = x[jI,a,b]
        # F_r[p,q] = x[a=q,b=p]: the fold adds alpha (x[b,a] - x[a,b]) at a<b
        be.packed_fold_virt(S, nv, nv, c2, self._off_rows, alpha=-0.5 * SQ2)
        return S

    def apply_pphh_pphh(self, U, S, beta, Ud=None):
        """S = beta*S + (pphh_pphh block) U   (matrix.py:136-143, 249-264)."""
        no, nc, nv = self.no, self.nc, self.nv
        # 2 A_ab(u f_vv) - f_oo u - f_cc u = (e_a + e_b - e_i - e_J) u for canonical orbitals
        be.mul(U, self.D0, out=S, beta=beta)
        if self.order_dd == 0:
            return S
        be.gemm(U, self.W, out=S, beta=1.0)                       # 1/2 u[iJcd] <ab||cd> = sum_{c<d}
        be.gemm(self.O, U.t(), out=S, beta=1.0)                   # <iJ||lK> u[lKab]  (N-major B)
        Ud = self._dense(U) if Ud is None else Ud
        U4 = Ud.view(no, nc, nv, nv)
        X1 = be.permute(U4, (0, 2, 1, 3)).reshape(no * nv, nc * nv)          # [(i,a),(K,c)] = u[iKac]
        T = be.gemm(X1, self.Ycv, alpha=-1.0)                                 # - u[iKac] <Kb||Jc>
        del X1
        X2 = be.permute(U4, (1, 2, 0, 3)).reshape(nc * nv, no * nv)          # [(J,b),(k,c)] = u[kJbc]
        be.gemm(self.Yov, X2, out=T, beta=1.0)                                # + <ic||ka> u[kJbc]
        del X2
        # -2 A_ab(..) + 2 A_ab(..): S[(i,J),AB] += T[i,a,J,b] - T[i,b,J,a]
        be.packed_fold_virt(S, nv, nc * nv, T, self._off_T, alpha=1.0)
        return S

    # ------------------------------------------------------------------ matvec
    def matvec(self, v):
        u1 = v["ph"]
        U = self._packed(v["pphh"])
        Ud = self._dense(U)
        s1 = self.apply_ph_ph(u1)
        be.axpby(1.0, self.apply_ph_pphh(U, Ud), 1.0, s1)
        S = be.empty((self.R, self.npv))
        self.apply_pphh_pphh(U, S, 0.0, Ud)
        self.apply_pphh_ph(u1, S)
        return AmplitudeVector(ph=self._wrap_ph(s1), pphh=self._wrap_pphh(S))

    def matvec_host(self, u1_host, U_host, s1_host, S_host, edge_rows=None):
        """sigma = M u between pinned host buffers (ph: [nc, nv]; pphh: packed [(i,J), AB])."""
        v = AmplitudeVector(ph=self._wrap_ph(be.upload(u1_host)), pphh=self._wrap_pphh(be.upload(U_host)))
        sig = self.matvec(v)
        be.download(sig.ph._contig(), s1_host)
        be.download(sig.pphh._contig(), S_host)
        return sig

    def block_apply(self, block, tensor):
        if block == "ph_ph":
            return self._wrap_ph(self.apply_ph_ph(tensor))
        if block == "ph_pphh":
            return self._wrap_ph(self.apply_ph_pphh(self._packed(tensor)))
        S = be.zeros((self.R, self.npv))
        if block == "pphh_ph":
            return self._wrap_pphh(self.apply_pphh_ph(tensor, S))
        if block == "pphh_pphh":
            return self._wrap_pphh(self.apply_pphh_pphh(self._packed(tensor), S, 0.0))
        raise ValueError(block)
