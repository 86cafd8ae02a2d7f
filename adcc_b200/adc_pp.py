"""PP-ADC secular matrix blocks as contraction tables.

The working equations are those of adcc/adc_pp/matrix.py (0th-3rd order ph/pphh blocks and
their CVS variants: :93-143, :192-304, :353-391, :490-545, :606-618; intermediates :933-1094,
:1548-1550; diagonals :109-124, :212-231).  Instead of Python closures over lazy tensors,
every block is a table of terms ``(coefficient, expression)`` in a tiny expression language:

    E(subscripts, *operands)       pairwise/multi-operand contraction -> DMMA GEMM(s)
    A(expr, (i,j), ...)            successive antisymmetrisations 1/2 (T - P_ij T)
    S(expr, ((i,j),(k,l)))         simultaneous symmetrisation

Operand names are resolved against the reference state (``ovov``, ``fvv`` ...), the MP ground
state (``t2``), the intermediates (``adc2_i1`` ...) and the trial vector (``u1``, ``u2``).
The interpreter (AdcMatrix) evaluates a block's terms and combines them in ONE fused
linear-combination launch; because the table is data, it can also pre-stage constant operands
in their GEMM layout at construction time and count the executed flops per matvec.
"""
from math import sqrt

from . import block as b
from .functions import einsum, direct_sum, zeros_like
from .AmplitudeVector import AmplitudeVector

SQ2 = sqrt(2.0)


def E(subs, *ops):
    return ("einsum", subs, ops)


def A(expr, *pairs):
    return ("asym", expr, pairs)


def S(expr, pairs):
    return ("sym", expr, pairs)


# ------------------------------------------------------------------------------ block tables
# key: (variant, "bra_ket", order) -> (input block, output block, [(coeff, expr), ...])
BLOCKS = {}


def _reg(variant, name, order, inb, outb, terms):
    BLOCKS[(variant, name, order)] = (inb, outb, terms)


# ---- singles-singles ---------------------------------------------------------- matrix.py:93-106,192-209
for _v, _fcc, _cvcv in (("", "foo", "ovov"), ("cvs", "fcc", "cvcv")):
    _reg(_v, "ph_ph", 0, "ph", "ph", [
        (+1.0, E("ib,ab->ia", "u1", "fvv")),
        (-1.0, E("IJ,Ja->Ia", _fcc, "u1")),
    ])
    _reg(_v, "ph_ph", 1, "ph", "ph", [
        (+1.0, E("ib,ab->ia", "u1", "fvv")),
        (-1.0, E("IJ,Ja->Ia", _fcc, "u1")),
        (-1.0, E("JaIb,Jb->Ia", _cvcv, "u1")),
    ])
# matrix.py:353-391
_reg("", "ph_ph", 2, "ph", "ph", [
    (+1.0, E("ib,ab->ia", "u1", "adc2_i1")),
    (-1.0, E("ij,ja->ia", "adc2_i2", "u1")),
    (-1.0, E("jaib,jb->ia", "ovov", "u1")),
    (-0.5, E("ikac,kc->ia", "term_t2_eri", "u1")),
])
_reg("cvs", "ph_ph", 2, "ph", "ph", [
    (+1.0, E("ib,ab->ia", "u1", "adc2_i1")),
    (-1.0, E("ij,ja->ia", "fcc", "u1")),
    (-1.0, E("JaIb,Jb->Ia", "cvcv", "u1")),
])
# matrix.py:606-618
_reg("", "ph_ph", 3, "ph", "ph", [(+1.0, E("iajb,jb->ia", "adc3_m11", "u1"))])
_reg("cvs", "ph_ph", 3, "ph", "ph", [(+1.0, E("iajb,jb->ia", "cvs_adc3_m11", "u1"))])

# ---- doubles-doubles ---------------------------------------------------------- matrix.py:127-143,234-264
_reg("", "pphh_pphh", 0, "pphh", "pphh", [
    (+2.0, A(E("ijac,bc->ijab", "u2", "fvv"), (2, 3))),
    (-2.0, A(E("ik,kjab->ijab", "foo", "u2"), (0, 1))),
])
_reg("cvs", "pphh_pphh", 0, "pphh", "pphh", [
    (+2.0, A(E("iJac,bc->iJab", "u2", "fvv"), (2, 3))),
    (-1.0, E("ik,kJab->iJab", "foo", "u2")),
    (-1.0, E("JK,iKab->iJab", "fcc", "u2")),
])
_reg("", "pphh_pphh", 1, "pphh", "pphh", [
    (+2.0, A(E("ijac,bc->ijab", "u2", "fvv"), (2, 3))),
    (-2.0, A(E("ik,kjab->ijab", "foo", "u2"), (0, 1))),
    (-4.0, A(E("ikac,kbjc->ijab", "u2", "ovov"), (0, 1), (2, 3))),
    (+0.5, E("ijkl,klab->ijab", "oooo", "u2")),
    (+0.5, E("ijcd,abcd->ijab", "u2", "vvvv")),
])
_reg("cvs", "pphh_pphh", 1, "pphh", "pphh", [
    (+2.0, A(E("iJac,bc->iJab", "u2", "fvv"), (2, 3))),
    (-1.0, E("ik,kJab->iJab", "foo", "u2")),
    (-1.0, E("JK,iKab->iJab", "fcc", "u2")),
    (-2.0, A(E("iKac,KbJc->iJab", "u2", "cvcv"), (2, 3))),
    (+2.0, A(E("icka,kJbc->iJab", "ovov", "u2"), (2, 3))),
    (+1.0, E("iJlK,lKab->iJab", "ococ", "u2")),
    (+0.5, E("iJcd,abcd->iJab", "u2", "vvvv")),
])

# ---- couplings ---------------------------------------------------------------- matrix.py:270-304,490-545
_reg("", "ph_pphh", 1, "pphh", "ph", [
    (+1.0, E("jkib,jkab->ia", "ooov", "u2")),
    (+1.0, E("ijbc,jabc->ia", "u2", "ovvv")),
])
_reg("cvs", "ph_pphh", 1, "pphh", "ph", [
    (+SQ2, E("jKIb,jKab->Ia", "occv", "u2")),
    (-1.0 / SQ2, E("jIbc,jabc->Ia", "u2", "ovvv")),
])
_reg("", "pphh_ph", 1, "ph", "pphh", [
    (+1.0, A(E("ic,jcab->ijab", "u1", "ovvv"), (0, 1))),
    (-1.0, A(E("ijka,kb->ijab", "ooov", "u1"), (2, 3))),
])
_reg("cvs", "pphh_ph", 1, "ph", "pphh", [
    (+SQ2, A(E("jIKb,Ka->jIab", "occv", "u1"), (2, 3))),
    (-1.0 / SQ2, E("Ic,jcab->jIab", "u1", "ovvv")),
])
_reg("", "ph_pphh", 2, "pphh", "ph", [
    (+1.0, E("jkib,jkab->ia", "adc3_pia", "u2")),
    (+1.0, E("ijbc,jabc->ia", "u2", "adc3_pib")),
    (+1.0, E("icab,cb->ia", "ovvv", E("jkcd,jkbd->cb", "u2", "t2"))),
    (+1.0, E("ijka,jk->ia", "ooov", E("jlbc,klbc->jk", "t2", "u2"))),
])
_reg("cvs", "ph_pphh", 2, "pphh", "ph", [
    (+2.0 / SQ2, E("lKIc,lKac->Ia", "cvs_adc3_pia", "u2")),
    (-1.0 / SQ2, E("lIcd,lacd->Ia", "u2", "adc3_pib")),
    (-1.0 / SQ2, E("jIKa,jK->Ia", "occv", E("ljcd,lKcd->jK", "t2", "u2"))),
])
_reg("", "pphh_ph", 2, "ph", "pphh", [
    (+1.0, A(E("ic,jcab->ijab", "u1", "adc3_pib"), (0, 1))),
    (+1.0, A(E("li,jlab->ijab", E("lkic,kc->li", "ooov", "u1"), "t2"), (0, 1))),
    (-1.0, A(E("ijka,kb->ijab", "adc3_pia", "u1"), (2, 3))),
    (-1.0, A(E("ijac,bc->ijab", "t2", E("kbcd,kd->bc", "ovvv", "u1")), (2, 3))),
])
_reg("cvs", "pphh_ph", 2, "ph", "pphh", [
    (-2.0 / SQ2, A(E("jIKa,Kb->jIab", "cvs_adc3_pia", "u1"), (2, 3))),
    (-1.0 / SQ2, E("Ic,jcab->jIab", "u1", "adc3_pib")),
    (-1.0 / SQ2, E("lI,jlab->jIab", E("lKIc,Kc->lI", "occv", "u1"), "t2")),
])
for _v in ("", "cvs"):
    for _name in ("ph_pphh", "pphh_ph"):
        _reg(_v, _name, 0, None, None, [])


# ------------------------------------------------------------------------------ intermediates
def _i_adc2_i1(hf, mp, im):
    return hf.fvv + 0.5 * einsum("ijac,ijbc->ab", mp.t2oo, hf.oovv).symmetrise()


def _i_adc2_i2(hf, mp, im):
    return hf.foo - 0.5 * einsum("ikab,jkab->ij", mp.t2oo, hf.oovv).symmetrise()


def _i_term_t2_eri(hf, mp, im):
    return (einsum("ijab,jkbc->ikac", mp.t2oo, hf.oovv)
            + einsum("ijab,jkbc->ikac", hf.oovv, mp.t2oo))


def _i_t2sq(hf, mp, im):
    return einsum("ikac,jkbc->iajb", mp.t2oo, mp.t2oo)


class DenseEriOps:
    """The contractions of adc3_i1 / adc3_m11 that touch the ovvv or vvvv block
    (adcc/adc_pp/matrix.py:951-953, 1014, 1029-1033), on dense blocks.  A reference state may carry
    its own implementation as ``hf.big_eri_ops`` (adcc_b200/packed_adc3.py: pair-packed operands,
    exchange-ordered slabs) for sizes at which the dense blocks cannot exist."""

    def __init__(self, hf):
        self.hf = hf

    def ovvv_iabc_ic(self, p_ov):
        return einsum("iabc,ic->ab", self.hf.ovvv, p_ov)

    def vvvv_acbd_cd(self, p_vv):
        return einsum("acbd,cd->ab", self.hf.vvvv, p_vv)

    def ovvv_ibac_jc(self, x_ov):
        return einsum("ibac,jc->iajb", self.hf.ovvv, x_ov)

    def m11_v4(self, ovov, t2, t2sq):
        return (0.5 * einsum("icjd,acbd->iajb", ovov, einsum("klac,klbd->acbd", t2, t2))
                - einsum("idjc,acbd->iajb", t2sq, self.hf.vvvv))


def _eri_ops(hf):
    return getattr(hf, "big_eri_ops", None) or DenseEriOps(hf)


def _adc3_i1(hf, mp, im):
    td2 = mp.td2(b.oovv)
    p0 = mp.second_order_dm_correction(apply_cvs=hf.has_core_occupied_space)
    t2eri_sum = (mp.t2eri(b.oovv, b.ov).transpose((1, 0, 2, 3)) - 0.25 * mp.t2eri(b.oovv, b.vv))
    ops = _eri_ops(hf)
    return ((0.5 * einsum("ijac,ijbc->ab", mp.t2oo + td2, hf.oovv)
             - 1.0 * einsum("ijac,ijcb->ab", mp.t2oo, t2eri_sum)
             - 2.0 * ops.ovvv_iabc_ic(p0.ov)).symmetrise()
            + einsum("iajb,ij->ab", hf.ovov, p0.oo)
            + ops.vvvv_acbd_cd(p0.vv))


def _adc3_i2(hf, mp, im):
    td2 = mp.td2(b.oovv)
    p0 = mp.mp2_diffdm
    t2eri_sum = mp.t2eri(b.oovv, b.ov) + 0.25 * mp.t2eri(b.oovv, b.oo)
    return ((0.5 * einsum("ikab,jkab->ij", mp.t2oo + td2, hf.oovv)
             - 1.0 * einsum("ikab,jkab->ij", mp.t2oo, t2eri_sum)
             + 2.0 * einsum("kija,ka->ij", hf.ooov, p0.ov)).symmetrise()
            - einsum("ikjl,kl->ij", hf.oooo, p0.oo)
            - einsum("iajb,ab->ij", hf.ovov, p0.vv))


def _cvs_adc3_i2(hf, mp, im):
    p0 = mp.second_order_dm_correction(apply_cvs=True)
    return (2.0 * einsum("kIJa,ka->IJ", hf.occv, p0.ov).symmetrise()
            - 1.0 * einsum("kIlJ,kl->IJ", hf.ococ, p0.oo)
            - 1.0 * einsum("IaJb,ab->IJ", hf.cvcv, p0.vv))


def _delta(t):
    d = zeros_like(t)
    d.set_mask("ii", 1.0)
    return d


def _i_adc3_m11(hf, mp, im):
    td2 = mp.td2(b.oovv)
    p0 = mp.mp2_diffdm
    i1, i2 = _adc3_i1(hf, mp, im), _adc3_i2(hf, mp, im)
    t2sq = im.t2sq
    t2 = mp.t2oo
    d_oo, d_vv = _delta(hf.foo), _delta(hf.fvv)
    t2eri_sum = (2.0 * mp.t2eri(b.oovv, b.ov).symmetrise((0, 1), (2, 3))
                 + 0.5 * mp.t2eri(b.oovv, b.vv) + 0.5 * mp.t2eri(b.oovv, b.oo))
    ov2 = 2.0 * p0.ov
    ops = _eri_ops(hf)
    inner = (einsum("jkbc,ikac->iajb", hf.oovv, t2 + td2)
             - einsum("jkbc,ikac->iajb", t2, t2eri_sum)
             - ops.ovvv_ibac_jc(ov2)
             - einsum("ikja,kb->iajb", hf.ooov, ov2)
             - einsum("jaic,bc->iajb", hf.ovov, p0.vv)
             + einsum("ik,jakb->iajb", p0.oo, hf.ovov)
             + einsum("ibkc,kajc->iajb", hf.ovov, 2.0 * t2sq))
    return (einsum("ij,ab->iajb", d_oo, hf.fvv + i1)
            - einsum("ij,ab->iajb", hf.foo - i2, d_vv)
            - einsum("jaib->iajb", hf.ovov)
            - inner.symmetrise((0, 2), (1, 3))
            + ops.m11_v4(hf.ovov, t2, t2sq)
            + 0.5 * einsum("ikjl,kalb->iajb", einsum("ikcd,jlcd->ikjl", t2, t2), hf.ovov)
            - einsum("iljk,kalb->iajb", hf.oooo, t2sq))


def _i_cvs_adc3_m11(hf, mp, im):
    i1, i2 = _adc3_i1(hf, mp, im), _cvs_adc3_i2(hf, mp, im)
    t2sq = im.t2sq
    t2 = mp.t2oo
    p0 = mp.second_order_dm_correction(apply_cvs=True)
    d_cc, d_vv = _delta(hf.fcc), _delta(hf.fvv)
    inner = (einsum("JaIc,bc->IaJb", hf.cvcv, p0.vv)
             - einsum("kIJa,kb->IaJb", hf.occv, 2.0 * p0.ov))
    return (einsum("IJ,ab->IaJb", d_cc, hf.fvv + i1)
            - einsum("IJ,ab->IaJb", hf.fcc - i2, d_vv)
            - einsum("JaIb->IaJb", hf.cvcv)
            + inner.symmetrise((0, 2), (1, 3))
            + 0.5 * einsum("IcJd,acbd->IaJb", hf.cvcv, einsum("klac,klbd->acbd", t2, t2))
            - einsum("lIkJ,kalb->IaJb", hf.ococ, t2sq))


def _i_adc3_pia(hf, mp, im):
    return (hf.ooov - 2.0 * mp.t2eri(b.ooov, b.ov).antisymmetrise(0, 1)
            - 0.5 * mp.t2eri(b.ooov, b.vv))


def _i_cvs_adc3_pia(hf, mp, im):
    return hf.occv - einsum("jlac,lKIc->jIKa", mp.t2oo, hf.occv)


def _i_adc3_pib(hf, mp, im):
    return (hf.ovvv + 2.0 * mp.t2eri(b.ovvv, b.ov).antisymmetrise(2, 3)
            - 0.5 * mp.t2eri(b.ovvv, b.oo))


INTERMEDIATES = {
    "adc2_i1": _i_adc2_i1, "adc2_i2": _i_adc2_i2, "term_t2_eri": _i_term_t2_eri,
    "t2sq": _i_t2sq, "adc3_m11": _i_adc3_m11, "cvs_adc3_m11": _i_cvs_adc3_m11,
    "adc3_pia": _i_adc3_pia, "cvs_adc3_pia": _i_cvs_adc3_pia, "adc3_pib": _i_adc3_pib,
}


# ------------------------------------------------------------------------------ diagonals
def diagonal(variant, name, order, hf, mp, im):
    """Diagonal contribution of a block (AmplitudeVector or 0)."""
    cvs = variant == "cvs"
    fCC = hf.fcc if cvs else hf.foo
    if name == "ph_ph":
        if order == 0:
            return AmplitudeVector(ph=direct_sum("a-i->ia", hf.fvv.diagonal(), fCC.diagonal()))
        if order == 1:
            CvCv = hf.cvcv if cvs else hf.ovov
            return AmplitudeVector(ph=(direct_sum("a-i->ia", hf.fvv.diagonal(), fCC.diagonal())
                                       - einsum("IaIa->Ia", CvCv)))
        if order == 2 and not cvs:
            return AmplitudeVector(ph=(
                direct_sum("a-i->ia", im.adc2_i1.diagonal(), im.adc2_i2.diagonal())
                - einsum("IaIa->Ia", hf.ovov)
                - einsum("ikac,ikac->ia", mp.t2oo, hf.oovv)))
        if order == 2 and cvs:
            return AmplitudeVector(ph=(direct_sum("a-i->ia", im.adc2_i1.diagonal(), hf.fcc.diagonal())
                                       - einsum("IaIa->Ia", hf.cvcv)))
        if order == 3:
            m11 = im.cvs_adc3_m11 if cvs else im.adc3_m11
            return AmplitudeVector(ph=einsum("iaia->ia", m11))
    if name == "pphh_pphh":
        if order == 0:      # matrix.py:109-124
            res = direct_sum("-i-J+a+b->iJab", hf.foo.diagonal(), fCC.diagonal(),
                             hf.fvv.diagonal(), hf.fvv.diagonal()).symmetrise(2, 3)
            if not cvs:
                res = res.symmetrise(0, 1)
            return AmplitudeVector(pphh=res)
        if order == 1:      # matrix.py:212-231
            d_ov = (direct_sum("a-i->ia", hf.fvv.diagonal(), hf.foo.diagonal())
                    - 2.0 * einsum("iaia->ia", hf.ovov))
            if cvs:
                d_Cv = (direct_sum("a-I->Ia", hf.fvv.diagonal(), hf.fcc.diagonal())
                        - 2.0 * einsum("IaIa->Ia", hf.cvcv))
                diag_oC = einsum("iJiJ->iJ", hf.ococ)
            else:
                d_Cv = d_ov
                diag_oC = einsum("ijij->ij", hf.oooo).symmetrise()
            diag_vv = einsum("abab->ab", hf.vvvv).symmetrise()
            return AmplitudeVector(pphh=(
                direct_sum("ia+Jb->iJab", d_ov, d_Cv).symmetrise(2, 3)
                + direct_sum("iJ+ab->iJab", diag_oC, diag_vv)))
    return 0
