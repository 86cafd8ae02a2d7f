"""Moller-Plesset ground state on the device (builds feeding the ADC matrix).

Mirrors adcc/GroundState.py:40-75,210-241,285-324 and adcc/LazyMp.py:37-151,245-296: ``df``,
``td1``/``t2``/``t2oo``/``t2oc``/``t2cc``, ``ts2``, ``td2``, ``t2eri`` (pi1..pi7),
``second_order_dm_correction``/``mp2_diffdm``, ``energy_correction``, ``energy``.  All
contractions run through the DMMA GEMM, denominators through the elementwise kernels; every
result is cached and immutable like the reference's ``cached_member_function`` results.
"""
from . import block as b
from .functions import einsum, direct_sum
from .MoSpaces import split_spaces
from .ReferenceState import ReferenceState
from .timings import Timer


class DensityBlocks(dict):
    """Minimal block container for the one-particle density corrections (attribute access
    ``p0.oo``, ``p0.ov``, ``p0.vv`` like adcc.OneParticleDensity)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class LazyMp:
    def __init__(self, hf):
        if not isinstance(hf, ReferenceState):
            hf = ReferenceState(hf)
        self.reference_state = hf
        self.mospaces = hf.mospaces
        self.timer = Timer()
        self.has_core_occupied_space = hf.has_core_occupied_space
        self._cache = {}

    def _cached(self, key, build):
        if key not in self._cache:
            with self.timer.record("/".join(str(k) for k in key)):
                val = build()
            if hasattr(val, "set_immutable"):
                val._contig()
                val.set_immutable()
            self._cache[key] = val
        return self._cache[key]

    # GroundState.py:210-216
    def df(self, space):
        s1, s2 = split_spaces(space)
        hf = self.reference_state
        return self._cached(("df", s1, s2), lambda: direct_sum(
            "-i+a->ia", hf.fock(s1 + s1).diagonal(), hf.fock(s2 + s2).diagonal()))

    # LazyMp.py:37-54
    def td1(self, space):
        sp = split_spaces(space)
        assert all(s == b.v for s in sp[2:])
        hf = self.reference_state

        def build():
            eia = self.df(sp[0] + b.v)
            ejb = self.df(sp[1] + b.v)
            return hf.eri("".join(sp)) / direct_sum("ia+jb->ijab", eia, ejb).symmetrise((2, 3))
        return self._cached(("td1",) + tuple(sp), build)

    t2 = td1

    @property
    def t2oo(self): return self.td1(b.oovv)
    @property
    def t2oc(self): return self.td1(b.ocvv)
    @property
    def t2cc(self): return self.td1(b.ccvv)

    # GroundState.py:219-241
    def t2eri(self, space, contraction):
        hf = self.reference_state
        table = {
            b.ooov + b.vv: ("ijbc,kabc->ijka", b.ovvv),
            b.ooov + b.ov: ("ilab,lkjb->ijka", b.ooov),
            b.oovv + b.oo: ("klab,ijkl->ijab", b.oooo),
            b.oovv + b.ov: ("jkac,kbic->ijab", b.ovov),
            b.oovv + b.vv: ("ijcd,abcd->ijab", b.vvvv),
            b.ovvv + b.oo: ("jkbc,jkia->iabc", b.ooov),
            b.ovvv + b.ov: ("ijbd,jcad->iabc", b.ovvv),
        }
        key = space + contraction
        if key not in table:
            raise NotImplementedError(f"t2eri intermediate not implemented for space '{space}' "
                                      f"and contraction '{contraction}'.")
        subs, blk = table[key]
        return self._cached(("t2eri", key), lambda: einsum(subs, self.t2oo, hf.eri(blk)))

    # LazyMp.py:56-133
    def ts2(self, space, apply_cvs=False):
        sp = split_spaces(space)
        assert len(sp) == 2 and sp[1] == b.v
        if sp[0] == b.o:
            return self.ts2_ov(apply_cvs=apply_cvs)
        if sp[0] == b.c:
            return self.ts2_cv(apply_cvs=apply_cvs)
        raise NotImplementedError(f"Second-order MP singles amplitudes not implemented for space '{space}'.")

    def ts2_ov(self, apply_cvs=False):
        if apply_cvs and not self.has_core_occupied_space:
            raise RuntimeError("Cannot apply the CVS approximation to a ground state build on top "
                               "of a HF reference state without a core space.")
        hf = self.reference_state

        def build():
            t2oo = self.t2oo
            res = (0.5 * einsum("jabc,ijbc->ia", hf.ovvv, t2oo)
                   + 0.5 * einsum("jkib,jkab->ia", hf.ooov, t2oo))
            if self.has_core_occupied_space and not apply_cvs:
                res = res + (1.0 * einsum("jKab,jKib->ia", self.t2oc, hf.ocov)
                             + 0.5 * einsum("JKab,ibJK->ia", self.t2cc, hf.ovcc)
                             + 0.5 * einsum("iJbc,Jabc->ia", self.t2oc, hf.cvvv))
            return res / (-1.0 * self.df(b.ov))
        return self._cached(("ts2_ov", apply_cvs), build)

    def ts2_cv(self, apply_cvs=False):
        assert self.has_core_occupied_space
        if apply_cvs:
            raise RuntimeError("The 'cv' block of the second-order MP singles amplitude vanishes "
                               "within the CVS approximation.")
        hf = self.reference_state

        def build():
            t2oo, t2oc, t2cc = self.t2oo, self.t2oc, self.t2cc
            res = (1.0 * einsum("jKab,jKIb->Ia", t2oc, hf.occv)
                   + 0.5 * einsum("JKab,JKIb->Ia", t2cc, hf.cccv)
                   + 0.5 * einsum("jkab,jkIb->Ia", t2oo, hf.oocv)
                   + 0.5 * einsum("IJbc,Jabc->Ia", t2cc, hf.cvvv)
                   - 0.5 * einsum("jIbc,jabc->Ia", t2oc, hf.ovvv))
            return res / (-1.0 * self.df(b.cv))
        return self._cached(("ts2_cv",), build)

    # LazyMp.py:136-151
    def td2(self, space):
        if split_spaces(space) != split_spaces(b.oovv):
            raise NotImplementedError(f"T^D_2 term not implemented for space {space}.")

        def build():
            t2erit = self.t2eri(b.oovv, b.ov).transpose((1, 0, 2, 3))
            denom = direct_sum("ia,jb->ijab", self.df(b.ov), self.df(b.ov)).symmetrise(0, 1)
            num = (4.0 * t2erit.antisymmetrise(2, 3).antisymmetrise(0, 1)
                   - 0.5 * self.t2eri(b.oovv, b.vv)
                   - 0.5 * self.t2eri(b.oovv, b.oo))
            return num / denom
        return self._cached(("td2",), build)

    # GroundState.py:285-324
    def second_order_dm_correction(self, apply_cvs=False):
        if apply_cvs and not self.has_core_occupied_space:
            raise RuntimeError("Cannot apply the CVS approximation to a ground state build on top "
                               "of a HF reference state without a core space.")

        def build():
            t2oo = self.t2oo
            ret = DensityBlocks()
            ret["oo"] = -0.5 * einsum("ikab,jkab->ij", t2oo, t2oo)
            ret["ov"] = self.ts2(b.ov, apply_cvs=apply_cvs)
            ret["vv"] = 0.5 * einsum("ijac,ijbc->ab", t2oo, t2oo)
            if self.has_core_occupied_space and not apply_cvs:
                t2oc, t2cc = self.t2oc, self.t2cc
                ret["oo"] = ret["oo"] + (-0.5) * einsum("iLab,jLab->ij", t2oc, t2oc)
                ret["vv"] = ret["vv"] + (0.5 * einsum("IJac,IJbc->ab", t2cc, t2cc)
                                         + 1.0 * einsum("kJac,kJbc->ab", t2oc, t2oc))
                ret["cc"] = -0.5 * (einsum("kIab,kJab->IJ", t2oc, t2oc)
                                    + einsum("LIab,LJab->IJ", t2cc, t2cc))
                ret["oc"] = -0.5 * (einsum("kIab,kjab->jI", t2oc, t2oo)
                                    + einsum("ILab,jLab->jI", t2cc, t2oc))
                ret["cv"] = self.ts2(b.cv, apply_cvs=apply_cvs)
            return ret
        return self._cached(("dm2", apply_cvs), build)

    @property
    def mp2_diffdm(self):
        return self.second_order_dm_correction(apply_cvs=False)

    def diffdm(self, level=2, apply_cvs=False):
        if level != 2:
            raise NotImplementedError("Only the second-order difference density is on the hot path.")
        return self.second_order_dm_correction(apply_cvs=apply_cvs)

    def dipole_moment(self, level=2):
        """Ground-state dipole moment with the density corrected up to ``level``
        (adcc/GroundState.py:178-189)."""
        from .properties import ground_state_dipole_moment
        return self._cached(("dipole_moment", level), lambda: ground_state_dipole_moment(self, level))

    # LazyMp.py:245-296
    def energy_correction(self, level=2):
        assert level >= 0
        if level < 2:
            return 0.0
        hf = self.reference_state
        cvs = self.has_core_occupied_space
        if level == 2 and not cvs:
            terms = [(1.0, hf.oovv, self.t2oo)]
        elif level == 2 and cvs:
            terms = [(1.0, hf.oovv, self.t2oo), (2.0, hf.ocvv, self.t2oc), (1.0, hf.ccvv, self.t2cc)]
        elif level == 3 and not cvs:
            terms = [(1.0, hf.oovv, self.td2(b.oovv))]
        else:
            method = "CVS-MP" if cvs else "MP"
            raise NotImplementedError(f"{method} energy correction for level {level} not implemented.")
        return sum(-0.25 * pref * eri.dot(t2) for pref, eri, t2 in terms)

    def energy(self, level=2):
        assert level >= 0
        ene = self.reference_state.energy_scf
        for il in range(2, level + 1):
            ene += self.energy_correction(il)
        return ene
