"""Oracle ADC(n) secular matrix: block applies, diagonals, intermediates, matvec.

Restates (test infrastructure, see oracle/__init__.py)
  * adcc/adc_pp/matrix.py:93-143,192-304,353-391,490-545,606-618  (ph/pphh blocks, CVS)
  * adcc/adc_pp/matrix.py:933-1094,1548-1550 (adc2_i1/i2, adc3_i1/i2, cvs_adc3_i2,
    adc3_m11, cvs_adc3_m11, adc3_pia, cvs_adc3_pia, adc3_pib, t2sq)
  * adcc/AdcMatrix.py:74-130 (block orders), :390-396 (matvec = sum of block applies),
    :431-472 (doubles symmetrisation recipe)
on dense spin-orbital arrays.  A vector is a dict {"ph": (o,v), "pphh": (o,o,v,v)}
(CVS: (c,v) and (o,c,v,v)).
"""
from math import sqrt
import numpy as np
from .mp import OracleMp, es, asym, sym, sym_pairs

BLOCK_ORDERS = {   # adcc/AdcMatrix.py:74-78,113-130
    "adc0": {"ph_ph": 0},
    "adc1": {"ph_ph": 1},
    "adc2": {"ph_ph": 2, "ph_pphh": 1, "pphh_ph": 1, "pphh_pphh": 0},
    "adc2x": {"ph_ph": 2, "ph_pphh": 1, "pphh_ph": 1, "pphh_pphh": 1},
    "adc3": {"ph_ph": 3, "ph_pphh": 2, "pphh_ph": 2, "pphh_pphh": 1},
}


def _dsum4(a, b, c, d):
    return a[:, None, None, None] + b[None, :, None, None] + c[None, None, :, None] + d[None, None, None, :]


class OracleAdcMatrix:
    def __init__(self, method, refstate_or_mp, unique_pairs=False):
        """unique_pairs: evaluate the oooo / vvvv terms of the first-order pphh_pphh block over the
        canonical index pairs only (i<j, a<b, c<d), the way libtensor works on the symmetry-unique
        blocks of an antisymmetric tensor (libadcc_src/TensorImpl/as_lt_symmetry.cc:37-55);
        same result, 8x fewer flops in the dominant term.  Used for the CPU timing baseline."""
        mp = refstate_or_mp if isinstance(refstate_or_mp, OracleMp) else OracleMp(refstate_or_mp)
        self.unique_pairs = unique_pairs
        self.mp = mp
        self.hf = mp.hf
        self.method = method
        self.cvs = method.startswith("cvs-")
        base = method[4:] if self.cvs else method
        if self.cvs != self.hf.has_core_occupied_space:
            raise ValueError("CVS method <-> core space mismatch (adc_pp/matrix.py:66-73)")
        self.block_orders = dict(BLOCK_ORDERS[base])
        self._im = {}
        self.blocks = {}
        diag = {}
        for blk, order in self.block_orders.items():
            name = "_block_" + ("cvs_" if self.cvs else "") + blk + "_" + str(order)
            apply, d = getattr(self, name)()
            self.blocks[blk] = apply
            if d is not None:
                for k, v in d.items():
                    diag[k] = diag.get(k, 0) + v
        self._diagonal = diag
        hf = self.hf
        occ = "o2" if self.cvs else "o1"
        self.shapes = {"ph": (hf.n_orbs[occ], hf.n_orbs["v1"])}
        if "pphh" in diag:
            self.shapes["pphh"] = (hf.n_orbs["o1"], hf.n_orbs[occ], hf.n_orbs["v1"], hf.n_orbs["v1"])
        self.axis_blocks = sorted(self.shapes, key=len)
        n = sum(int(np.prod(s)) for s in self.shapes.values())
        self.shape = (n, n)

    def diagonal(self):
        return self._diagonal

    def matvec(self, v):
        out = {}
        for blk, apply in self.blocks.items():
            res = apply(v)
            if res is None:
                continue
            for k, t in res.items():
                out[k] = out[k] + t if k in out else t
        return out

    def block_apply(self, block, tensor):
        outb, inb = block.split("_")
        return self.blocks[block]({inb: tensor})[outb]

    def symmetrise_doubles(self, u2):
        """AdcMatrix.py:431-472."""
        if self.cvs:
            return asym(u2, 2, 3)
        return sym_pairs(asym(asym(u2, 0, 1), 2, 3), (0, 1), (2, 3))

    # ------------------------------------------------------------------ intermediates
    def _get(self, name):
        if name not in self._im:
            self._im[name] = getattr(self, "_im_" + name)()
        return self._im[name]

    def _im_adc2_i1(self):
        hf, mp = self.hf, self.mp
        x = es("ijac,ijbc->ab", mp.t2oo, hf.oovv)
        return hf.fvv + 0.5 * sym(x, 0, 1)

    def _im_adc2_i2(self):
        hf, mp = self.hf, self.mp
        x = es("ikab,jkab->ij", mp.t2oo, hf.oovv)
        return hf.foo - 0.5 * sym(x, 0, 1)

    def _im_t2sq(self):
        return es("ikac,jkbc->iajb", self.mp.t2oo, self.mp.t2oo)

    def _adc3_i1(self):
        hf, mp = self.hf, self.mp
        td2 = mp.td2()
        p0 = mp.mp2_diffdm(apply_cvs=hf.has_core_occupied_space)
        t2eri_sum = (es("jicb->ijcb", mp.t2eri("oovv", "ov")) - 0.25 * mp.t2eri("oovv", "vv"))
        return (sym(0.5 * es("ijac,ijbc->ab", mp.t2oo + td2, hf.oovv)
                    - 1.0 * es("ijac,ijcb->ab", mp.t2oo, t2eri_sum)
                    - 2.0 * es("iabc,ic->ab", hf.ovvv, p0["ov"]), 0, 1)
                + es("iajb,ij->ab", hf.ovov, p0["oo"])
                + es("acbd,cd->ab", hf.vvvv, p0["vv"]))

    def _adc3_i2(self):
        hf, mp = self.hf, self.mp
        td2 = mp.td2()
        p0 = mp.mp2_diffdm()
        t2eri_sum = mp.t2eri("oovv", "ov") + 0.25 * mp.t2eri("oovv", "oo")
        return (sym(0.5 * es("ikab,jkab->ij", mp.t2oo + td2, hf.oovv)
                    - 1.0 * es("ikab,jkab->ij", mp.t2oo, t2eri_sum)
                    + 2.0 * es("kija,ka->ij", hf.ooov, p0["ov"]), 0, 1)
                - es("ikjl,kl->ij", hf.oooo, p0["oo"])
                - es("iajb,ab->ij", hf.ovov, p0["vv"]))

    def _cvs_adc3_i2(self):
        hf, mp = self.hf, self.mp
        p0 = mp.mp2_diffdm(apply_cvs=True)
        return (2.0 * sym(es("kIJa,ka->IJ", hf.eri("occv"), p0["ov"]), 0, 1)
                - 1.0 * es("kIlJ,kl->IJ", hf.eri("ococ"), p0["oo"])
                - 1.0 * es("IaJb,ab->IJ", hf.eri("cvcv"), p0["vv"]))

    def _im_adc3_m11(self):
        hf, mp = self.hf, self.mp
        td2 = mp.td2()
        p0 = mp.mp2_diffdm()
        i1, i2 = self._adc3_i1(), self._adc3_i2()
        t2sq = self._get("t2sq")
        d_oo = np.eye(hf.n_orbs["o1"])
        d_vv = np.eye(hf.n_orbs["v1"])
        t2eri_sum = (2.0 * sym_pairs(mp.t2eri("oovv", "ov"), (0, 1), (2, 3))
                     + 0.5 * mp.t2eri("oovv", "vv") + 0.5 * mp.t2eri("oovv", "oo"))
        t2 = mp.t2oo
        inner = (es("jkbc,ikac->iajb", hf.oovv, t2 + td2)
                 - es("jkbc,ikac->iajb", t2, t2eri_sum)
                 - es("ibac,jc->iajb", hf.ovvv, 2.0 * p0["ov"])
                 - es("ikja,kb->iajb", hf.ooov, 2.0 * p0["ov"])
                 - es("jaic,bc->iajb", hf.ovov, p0["vv"])
                 + es("ik,jakb->iajb", p0["oo"], hf.ovov)
                 + es("ibkc,kajc->iajb", hf.ovov, 2.0 * t2sq))
        return (es("ij,ab->iajb", d_oo, hf.fvv + i1)
                - es("ij,ab->iajb", hf.foo - i2, d_vv)
                - es("jaib->iajb", hf.ovov)
                - sym_pairs(inner, (0, 2), (1, 3))
                + 0.5 * es("icjd,acbd->iajb", hf.ovov, es("klac,klbd->acbd", t2, t2))
                + 0.5 * es("ikjl,kalb->iajb", es("ikcd,jlcd->ikjl", t2, t2), hf.ovov)
                - es("iljk,kalb->iajb", hf.oooo, t2sq)
                - es("idjc,acbd->iajb", t2sq, hf.vvvv))

    def _im_cvs_adc3_m11(self):
        hf, mp = self.hf, self.mp
        i1, i2 = self._adc3_i1(), self._cvs_adc3_i2()
        t2sq = self._get("t2sq")
        p0 = mp.mp2_diffdm(apply_cvs=True)
        d_cc = np.eye(hf.n_orbs["o2"])
        d_vv = np.eye(hf.n_orbs["v1"])
        t2 = mp.t2oo
        cvcv, occv = hf.eri("cvcv"), hf.eri("occv")
        inner = (es("JaIc,bc->IaJb", cvcv, p0["vv"]) - es("kIJa,kb->IaJb", occv, 2.0 * p0["ov"]))
        return (es("IJ,ab->IaJb", d_cc, hf.fvv + i1)
                - es("IJ,ab->IaJb", hf.fcc - i2, d_vv)
                - es("JaIb->IaJb", cvcv)
                + sym_pairs(inner, (0, 2), (1, 3))
                + 0.5 * es("IcJd,acbd->IaJb", cvcv, es("klac,klbd->acbd", t2, t2))
                - es("lIkJ,kalb->IaJb", hf.eri("ococ"), t2sq))

    def _im_adc3_pia(self):
        hf, mp = self.hf, self.mp
        return (hf.ooov - 2.0 * asym(mp.t2eri("ooov", "ov"), 0, 1) - 0.5 * mp.t2eri("ooov", "vv"))

    def _im_cvs_adc3_pia(self):
        hf, mp = self.hf, self.mp
        occv = hf.eri("occv")
        return occv - es("jlac,lKIc->jIKa", mp.t2oo, occv)

    def _im_adc3_pib(self):
        hf, mp = self.hf, self.mp
        return (hf.ovvv + 2.0 * asym(mp.t2eri("ovvv", "ov"), 2, 3) - 0.5 * mp.t2eri("ovvv", "oo"))

    # ------------------------------------------------------------------ singles blocks
    def _fCC(self):
        return self.hf.fcc if self.cvs else self.hf.foo

    def _block_ph_ph_0(self):
        hf, fCC = self.hf, self._fCC()
        diag = {"ph": np.diag(hf.fvv)[None, :] - np.diag(fCC)[:, None]}

        def apply(v):
            u = v["ph"]
            return {"ph": es("ib,ab->ia", u, hf.fvv) - es("IJ,Ja->Ia", fCC, u)}
        return apply, diag
    _block_cvs_ph_ph_0 = _block_ph_ph_0

    def _block_ph_ph_1(self):
        hf, fCC = self.hf, self._fCC()
        CvCv = hf.eri("cvcv") if self.cvs else hf.ovov
        diag = {"ph": np.diag(hf.fvv)[None, :] - np.diag(fCC)[:, None] - es("IaIa->Ia", CvCv)}

        def apply(v):
            u = v["ph"]
            return {"ph": (es("ib,ab->ia", u, hf.fvv) - es("IJ,Ja->Ia", fCC, u)
                           - es("JaIb,Jb->Ia", CvCv, u))}
        return apply, diag
    _block_cvs_ph_ph_1 = _block_ph_ph_1

    def _block_ph_ph_2(self):
        hf, mp = self.hf, self.mp
        i1, i2 = self._get("adc2_i1"), self._get("adc2_i2")
        diag = {"ph": (np.diag(i1)[None, :] - np.diag(i2)[:, None]
                       - es("IaIa->Ia", hf.ovov) - es("ikac,ikac->ia", mp.t2oo, hf.oovv))}
        term_t2_eri = (es("ijab,jkbc->ikac", mp.t2oo, hf.oovv)
                       + es("ijab,jkbc->ikac", hf.oovv, mp.t2oo))

        def apply(v):
            u = v["ph"]
            return {"ph": (es("ib,ab->ia", u, i1) - es("ij,ja->ia", i2, u)
                           - es("jaib,jb->ia", hf.ovov, u)
                           - 0.5 * es("ikac,kc->ia", term_t2_eri, u))}
        return apply, diag

    def _block_cvs_ph_ph_2(self):
        hf = self.hf
        i1 = self._get("adc2_i1")
        cvcv = hf.eri("cvcv")
        diag = {"ph": np.diag(i1)[None, :] - np.diag(hf.fcc)[:, None] - es("IaIa->Ia", cvcv)}

        def apply(v):
            u = v["ph"]
            return {"ph": (es("ib,ab->ia", u, i1) - es("ij,ja->ia", hf.fcc, u)
                           - es("JaIb,Jb->Ia", cvcv, u))}
        return apply, diag

    def _block_ph_ph_3(self):
        m11 = self._get("cvs_adc3_m11" if self.cvs else "adc3_m11")
        diag = {"ph": es("iaia->ia", m11)}

        def apply(v):
            return {"ph": es("iajb,jb->ia", m11, v["ph"])}
        return apply, diag
    _block_cvs_ph_ph_3 = _block_ph_ph_3

    # ------------------------------------------------------------------ doubles blocks
    def _diag_pphh_0(self):
        hf, fCC = self.hf, self._fCC()
        res = sym(_dsum4(-np.diag(hf.foo), -np.diag(fCC), np.diag(hf.fvv), np.diag(hf.fvv)), 2, 3)
        if not self.cvs:
            res = sym(res, 0, 1)
        return {"pphh": res}

    def _diag_pphh_1(self):
        hf = self.hf
        d_ov = (np.diag(hf.fvv)[None, :] - np.diag(hf.foo)[:, None] - 2.0 * es("iaia->ia", hf.ovov))
        if self.cvs:
            d_Cv = (np.diag(hf.fvv)[None, :] - np.diag(hf.fcc)[:, None]
                    - 2.0 * es("IaIa->Ia", hf.eri("cvcv")))
            diag_oC = es("iJiJ->iJ", hf.eri("ococ"))
        else:
            d_Cv = d_ov
            diag_oC = sym(es("ijij->ij", hf.oooo), 0, 1)
        diag_vv = sym(es("abab->ab", hf.vvvv), 0, 1)
        return {"pphh": (sym(d_ov[:, None, :, None] + d_Cv[None, :, None, :], 2, 3)
                         + diag_oC[:, :, None, None] + diag_vv[None, None, :, :])}

    def _block_pphh_pphh_0(self):
        hf = self.hf

        def apply(v):
            u = v["pphh"]
            return {"pphh": (2 * asym(es("ijac,bc->ijab", u, hf.fvv), 2, 3)
                             - 2 * asym(es("ik,kjab->ijab", hf.foo, u), 0, 1))}
        return apply, self._diag_pphh_0()

    def _block_cvs_pphh_pphh_0(self):
        hf = self.hf

        def apply(v):
            u = v["pphh"]
            return {"pphh": (2 * asym(es("iJac,bc->iJab", u, hf.fvv), 2, 3)
                             - es("ik,kJab->iJab", hf.foo, u)
                             - es("JK,iKab->iJab", hf.fcc, u))}
        return apply, self._diag_pphh_0()

    def _pair_terms(self, u):
        """1/2 <ij||kl> u[klab] + 1/2 u[ijcd] <ab||cd> as two GEMMs over canonical pairs."""
        hf = self.hf
        no, nv = u.shape[0], u.shape[2]
        if not hasattr(self, "_pairs"):
            i, j = np.triu_indices(no, 1)
            a, b = np.triu_indices(nv, 1)
            W = np.ascontiguousarray(hf.vvvv[a[:, None], b[:, None], a[None, :], b[None, :]])
            O = np.ascontiguousarray(hf.oooo[i[:, None], j[:, None], i[None, :], j[None, :]])
            self._pairs = (i, j, a, b, W, O)
        i, j, a, b, W, O = self._pairs
        U = u[i[:, None], j[:, None], a[None, :], b[None, :]]          # [IJ, AB]
        S = U @ W.T + O @ U
        out = np.zeros_like(u)
        out[i[:, None], j[:, None], a[None, :], b[None, :]] = S
        out[j[:, None], i[:, None], a[None, :], b[None, :]] = -S
        out[i[:, None], j[:, None], b[None, :], a[None, :]] = -S
        out[j[:, None], i[:, None], b[None, :], a[None, :]] = S
        return out

    def _block_pphh_pphh_1(self):
        hf = self.hf

        def apply(v):
            u = v["pphh"]
            if self.unique_pairs:
                return {"pphh": (2 * asym(es("ijac,bc->ijab", u, hf.fvv), 2, 3)
                                 - 2 * asym(es("ik,kjab->ijab", hf.foo, u), 0, 1)
                                 + asym(asym(-4 * es("ikac,kbjc->ijab", u, hf.ovov), 0, 1), 2, 3)
                                 + self._pair_terms(u))}
            return {"pphh": (2 * asym(es("ijac,bc->ijab", u, hf.fvv), 2, 3)
                             - 2 * asym(es("ik,kjab->ijab", hf.foo, u), 0, 1)
                             + asym(asym(-4 * es("ikac,kbjc->ijab", u, hf.ovov), 0, 1), 2, 3)
                             + 0.5 * es("ijkl,klab->ijab", hf.oooo, u)
                             + 0.5 * es("ijcd,abcd->ijab", u, hf.vvvv))}
        return apply, self._diag_pphh_1()

    def _block_cvs_pphh_pphh_1(self):
        hf = self.hf
        cvcv, ococ = hf.eri("cvcv"), hf.eri("ococ")

        def apply(v):
            u = v["pphh"]
            return {"pphh": (2.0 * asym(es("iJac,bc->iJab", u, hf.fvv), 2, 3)
                             - 1.0 * es("ik,kJab->iJab", hf.foo, u)
                             - 1.0 * es("JK,iKab->iJab", hf.fcc, u)
                             + asym(-2.0 * es("iKac,KbJc->iJab", u, cvcv)
                                    + 2.0 * es("icka,kJbc->iJab", hf.ovov, u), 2, 3)
                             + 1.0 * es("iJlK,lKab->iJab", ococ, u)
                             + 0.5 * es("iJcd,abcd->iJab", u, hf.vvvv))}
        return apply, self._diag_pphh_1()

    # ------------------------------------------------------------------ couplings
    def _block_ph_pphh_1(self):
        hf = self.hf

        def apply(v):
            if "pphh" not in v:
                return None
            u = v["pphh"]
            return {"ph": es("jkib,jkab->ia", hf.ooov, u) + es("ijbc,jabc->ia", u, hf.ovvv)}
        return apply, None

    def _block_cvs_ph_pphh_1(self):
        hf = self.hf
        occv = hf.eri("occv")

        def apply(v):
            if "pphh" not in v:
                return None
            u = v["pphh"]
            return {"ph": (sqrt(2) * es("jKIb,jKab->Ia", occv, u)
                           - 1 / sqrt(2) * es("jIbc,jabc->Ia", u, hf.ovvv))}
        return apply, None

    def _block_pphh_ph_1(self):
        hf = self.hf

        def apply(v):
            if "ph" not in v:
                return None
            u = v["ph"]
            return {"pphh": (asym(es("ic,jcab->ijab", u, hf.ovvv), 0, 1)
                             - asym(es("ijka,kb->ijab", hf.ooov, u), 2, 3))}
        return apply, None

    def _block_cvs_pphh_ph_1(self):
        hf = self.hf
        occv = hf.eri("occv")

        def apply(v):
            if "ph" not in v:
                return None
            u = v["ph"]
            return {"pphh": (sqrt(2) * asym(es("jIKb,Ka->jIab", occv, u), 2, 3)
                             - 1 / sqrt(2) * es("Ic,jcab->jIab", u, hf.ovvv))}
        return apply, None

    def _block_ph_pphh_2(self):
        hf, mp = self.hf, self.mp
        pia, pib = self._get("adc3_pia"), self._get("adc3_pib")

        def apply(v):
            if "pphh" not in v:
                return None
            u = v["pphh"]
            return {"ph": (es("jkib,jkab->ia", pia, u) + es("ijbc,jabc->ia", u, pib)
                           + es("icab,jkcd,jkbd->ia", hf.ovvv, u, mp.t2oo)
                           + es("ijka,jlbc,klbc->ia", hf.ooov, mp.t2oo, u))}
        return apply, None

    def _block_cvs_ph_pphh_2(self):
        hf, mp = self.hf, self.mp
        pia, pib = self._get("cvs_adc3_pia"), self._get("adc3_pib")
        occv = hf.eri("occv")

        def apply(v):
            if "pphh" not in v:
                return None
            u = v["pphh"]
            return {"ph": (1 / sqrt(2)) * (2.0 * es("lKIc,lKac->Ia", pia, u)
                                           - es("lIcd,lacd->Ia", u, pib)
                                           - es("jIKa,ljcd,lKcd->Ia", occv, mp.t2oo, u))}
        return apply, None

    def _block_pphh_ph_2(self):
        hf, mp = self.hf, self.mp
        pia, pib = self._get("adc3_pia"), self._get("adc3_pib")

        def apply(v):
            if "ph" not in v:
                return None
            u = v["ph"]
            return {"pphh": (asym(es("ic,jcab->ijab", u, pib)
                                  + es("lkic,kc,jlab->ijab", hf.ooov, u, mp.t2oo), 0, 1)
                             + asym(-es("ijka,kb->ijab", pia, u)
                                    - es("ijac,kbcd,kd->ijab", mp.t2oo, hf.ovvv, u), 2, 3))}
        return apply, None

    def _block_cvs_pphh_ph_2(self):
        hf, mp = self.hf, self.mp
        pia, pib = self._get("cvs_adc3_pia"), self._get("adc3_pib")
        occv = hf.eri("occv")

        def apply(v):
            if "ph" not in v:
                return None
            u = v["ph"]
            return {"pphh": (1 / sqrt(2)) * (-2.0 * asym(es("jIKa,Kb->jIab", pia, u), 2, 3)
                                             - es("Ic,jcab->jIab", u, pib)
                                             - es("lKIc,Kc,jlab->jIab", occv, u, mp.t2oo))}
        return apply, None


def dominant_terms_sample(no, nv, ab_slab, jb_slab, reps=2):
    """Bounded sample of the two dominant contractions of the first-order pphh_pphh block AT FULL
    operand extents, for the CPU timing baseline (test infrastructure, never the product path):
    a column slab of  S[IJ,AB] += sum_CD U[IJ,CD] W[AB,CD]  (vvvv term over canonical pairs, as in
    ``_pair_terms``) and of  T[(i,a),(j,b)] = sum_kc X[(i,a),(k,c)] Y[(j,b),(k,c)]  (ovov term,
    ``ikac,kbjc->ijab`` of _block_pphh_pphh_1).  Operand values do not matter for a DGEMM timing.
    Returns (flops, seconds) of the best of ``reps`` repetitions."""
    import time
    npo, npv = no * (no - 1) // 2, nv * (nv - 1) // 2
    ab_slab, jb_slab = min(ab_slab, npv), min(jb_slab, no * nv)
    U = np.full((npo, npv), 0.5)
    W = np.full((ab_slab, npv), 0.25)
    X = np.full((no * nv, no * nv), 0.5)
    Y = np.full((jb_slab, no * nv), 0.25)
    flops = 2.0 * npo * ab_slab * npv + 2.0 * (no * nv) * jb_slab * (no * nv)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        S = U @ W.T
        T = X @ Y.T
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert S.shape == (npo, ab_slab) and T.shape == (no * nv, jb_slab)
    return flops, best
