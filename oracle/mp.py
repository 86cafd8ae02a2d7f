"""Oracle Moller-Plesset ground state (test infrastructure, see oracle/__init__.py).

Restates adcc/GroundState.py:210-241,285-324 (df, t2eri pi1..pi7, second-order density
correction) and adcc/LazyMp.py:37-151,245-269 (td1 = t2, ts2_ov/ts2_cv, td2, energy
corrections) on dense spin-orbital arrays.  Conventions: positive-denominator t2
(LazyMp.py:42-54); symmetrise/antisymmetrise include the 1/2 (libadcc_src/Tensor.hh:229-244).
"""
import numpy as np
from .refstate import split_space


def es(subs, *ops):
    return np.einsum(subs, *ops, optimize=True)


def swap(t, i, j):
    return np.swapaxes(t, i, j)


def asym(t, i, j):
    """Tensor.antisymmetrise(i, j): 1/2 (T - P_ij T)."""
    return 0.5 * (t - swap(t, i, j))


def sym(t, i, j):
    """Tensor.symmetrise(i, j): 1/2 (T + P_ij T)."""
    return 0.5 * (t + swap(t, i, j))


def sym_pairs(t, p1, p2):
    """Tensor.symmetrise((a,b),(c,d)): both transpositions applied simultaneously."""
    return 0.5 * (t + swap(swap(t, p1[0], p1[1]), p2[0], p2[1]))


class OracleMp:
    def __init__(self, refstate):
        self.hf = refstate
        self.has_core_occupied_space = refstate.has_core_occupied_space
        self._c = {}

    def _cached(self, key, fn):
        if key not in self._c:
            self._c[key] = fn()
        return self._c[key]

    def df(self, space):
        """GroundState.py:210-216: df[ia] = e_a - e_i."""
        s1, s2 = split_space(space)
        return self._cached(("df", s1, s2), lambda: (
            -np.diag(self.hf.fock(s1 + s1))[:, None] + np.diag(self.hf.fock(s2 + s2))[None, :]))

    def t2(self, space):
        """LazyMp.py:37-54: eri / sym_{2,3}(df_ia + df_jb)."""
        sp = split_space(space)

        def build():
            eia = self.df(sp[0] + sp[2])
            ejb = self.df(sp[1] + sp[3])
            denom = sym(eia[:, None, :, None] + ejb[None, :, None, :], 2, 3)
            return self.hf.eri("".join(sp)) / denom
        return self._cached(("t2",) + tuple(sp), build)

    td1 = t2

    @property
    def t2oo(self):
        return self.t2("oovv")

    @property
    def t2oc(self):
        return self.t2("ocvv")

    @property
    def t2cc(self):
        return self.t2("ccvv")

    def t2eri(self, space, contraction):
        """GroundState.py:219-241 (pi1..pi7)."""
        table = {
            ("ooov", "vv"): ("ijbc,kabc->ijka", "ovvv"),
            ("ooov", "ov"): ("ilab,lkjb->ijka", "ooov"),
            ("oovv", "oo"): ("klab,ijkl->ijab", "oooo"),
            ("oovv", "ov"): ("jkac,kbic->ijab", "ovov"),
            ("oovv", "vv"): ("ijcd,abcd->ijab", "vvvv"),
            ("ovvv", "oo"): ("jkbc,jkia->iabc", "ooov"),
            ("ovvv", "ov"): ("ijbd,jcad->iabc", "ovvv"),
        }
        subs, blk = table[(space, contraction)]
        return self._cached(("t2eri", space, contraction),
                            lambda: es(subs, self.t2oo, self.hf.eri(blk)))

    def ts2_ov(self, apply_cvs=False):
        """LazyMp.py:72-104."""
        def build():
            hf = self.hf
            t2oo = self.t2oo
            res = (0.5 * es("jabc,ijbc->ia", hf.ovvv, t2oo)
                   + 0.5 * es("jkib,jkab->ia", hf.ooov, t2oo))
            if self.has_core_occupied_space and not apply_cvs:
                res = res + (1.0 * es("jKab,jKib->ia", self.t2oc, hf.eri("ocov"))
                             + 0.5 * es("JKab,ibJK->ia", self.t2cc, hf.eri("ovcc"))
                             + 0.5 * es("iJbc,Jabc->ia", self.t2oc, hf.eri("cvvv")))
            return res / (-self.df("ov"))
        return self._cached(("ts2_ov", apply_cvs), build)

    def ts2_cv(self):
        """LazyMp.py:106-133."""
        def build():
            hf = self.hf
            res = (1.0 * es("jKab,jKIb->Ia", self.t2oc, hf.eri("occv"))
                   + 0.5 * es("JKab,JKIb->Ia", self.t2cc, hf.eri("cccv"))
                   + 0.5 * es("jkab,jkIb->Ia", self.t2oo, hf.eri("oocv"))
                   + 0.5 * es("IJbc,Jabc->Ia", self.t2cc, hf.eri("cvvv"))
                   - 0.5 * es("jIbc,jabc->Ia", self.t2oc, hf.ovvv))
            return res / (-self.df("cv"))
        return self._cached(("ts2_cv",), build)

    def td2(self, space="oovv"):
        """LazyMp.py:136-151."""
        def build():
            t2erit = self.t2eri("oovv", "ov").transpose(1, 0, 2, 3)
            dov = self.df("ov")
            denom = sym(dov[:, None, :, None] + dov[None, :, None, :], 0, 1)
            return (4.0 * asym(asym(t2erit, 2, 3), 0, 1)
                    - 0.5 * self.t2eri("oovv", "vv")
                    - 0.5 * self.t2eri("oovv", "oo")) / denom
        return self._cached(("td2",), build)

    def mp2_diffdm(self, apply_cvs=False):
        """GroundState.py:285-324: dict of blocks oo, ov, vv (+ cc, oc, cv)."""
        def build():
            t2oo = self.t2oo
            ret = {
                "oo": -0.5 * es("ikab,jkab->ij", t2oo, t2oo),
                "ov": self.ts2_ov(apply_cvs=apply_cvs),
                "vv": 0.5 * es("ijac,ijbc->ab", t2oo, t2oo),
            }
            if self.has_core_occupied_space and not apply_cvs:
                t2oc, t2cc = self.t2oc, self.t2cc
                ret["oo"] = ret["oo"] - 0.5 * es("iLab,jLab->ij", t2oc, t2oc)
                ret["vv"] = ret["vv"] + (0.5 * es("IJac,IJbc->ab", t2cc, t2cc)
                                         + 1.0 * es("kJac,kJbc->ab", t2oc, t2oc))
                ret["cc"] = -0.5 * (es("kIab,kJab->IJ", t2oc, t2oc)
                                    + es("LIab,LJab->IJ", t2cc, t2cc))
                ret["oc"] = -0.5 * (es("kIab,kjab->jI", t2oc, t2oo)
                                    + es("ILab,jLab->jI", t2cc, t2oc))
                ret["cv"] = self.ts2_cv()
            return ret
        return self._cached(("diffdm", apply_cvs), build)

    def energy_correction(self, level=2):
        """LazyMp.py:245-269."""
        hf = self.hf
        if level < 2:
            return 0.0
        cvs = self.has_core_occupied_space
        if level == 2 and not cvs:
            terms = [(1.0, hf.oovv, self.t2oo)]
        elif level == 2 and cvs:
            terms = [(1.0, hf.oovv, self.t2oo), (2.0, hf.eri("ocvv"), self.t2oc),
                     (1.0, hf.eri("ccvv"), self.t2cc)]
        elif level == 3 and not cvs:
            terms = [(1.0, hf.oovv, self.td2())]
        else:
            raise NotImplementedError
        return sum(-0.25 * p * float(np.vdot(e, t)) for p, e, t in terms)
