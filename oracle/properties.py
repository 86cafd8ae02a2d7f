"""Oracle transition densities / oscillator strengths / state densities (test infrastructure, see
oracle/__init__.py).

Restates adcc/adc_pp/transition_dm.py:36-97 (tdm_isr0/1/2, tdm_cvs_isr2),
adcc/OperatorIntegrals.py:37-76 (AO->MO, spin-block-diagonal) and
adcc/ElectronicTransition.py:178-183 (f = 2/3 |mu|^2 omega) on dense arrays.  State difference densities and
state dipole moments restate adcc/adc_pp/state_diffdm.py:37-160, adcc/ElectronicStates.py:247-264 and
adcc/GroundState.py:137-146, 178-189.  Pinned by the goldens the reference's own property code
produced (tests/golden/reference_adcc_h2o_sto3g.npz, tests/reference_golden_common.py).
"""
from math import sqrt
import numpy as np
from .mp import es


def transition_dm(matrix, vec, level=None):
    """level: ISR order of the density (default: the method's own, at most 2; adcc.cis uses 0)."""
    mp, cvs = matrix.mp, matrix.cvs
    base = matrix.method[4:] if cvs else matrix.method
    if level is None:
        level = min(int(base[3]), 2)
    u1 = vec["ph"]
    C = "c" if cvs else "o"
    dm = {"v" + C: u1.T.copy()}
    if cvs:
        if level >= 2:
            u2, t2 = vec["pphh"], mp.t2oo
            p0 = mp.mp2_diffdm(apply_cvs=True)
            dm["oc"] = -es("ja,Ia->jI", p0["ov"], u1) + (1 / sqrt(2)) * es("kIab,jkab->jI", u2, t2)
            dm["vc"] = dm["vc"] - 0.5 * es("ab,Ib->aI", p0["vv"], u1)
        return dm
    if level >= 1:
        dm["ov"] = -es("ijab,jb->ia", mp.t2oo, u1)
    if level >= 2:
        u2, t2, td2 = vec["pphh"], mp.t2oo, mp.td2()
        p0 = mp.mp2_diffdm()
        dm["oo"] = -es("ia,ja->ij", p0["ov"], u1) - es("ikab,jkab->ji", u2, t2)
        dm["vv"] = es("ia,ib->ab", u1, p0["ov"]) + es("ijac,ijbc->ab", u2, t2)
        dm["ov"] = dm["ov"] - es("ijab,jb->ia", td2, u1)
        dm["vo"] = dm["vo"] + 0.5 * (es("ijab,jkbc,kc->ai", t2, t2, u1)
                                     - es("ab,ib->ai", p0["vv"], u1)
                                     + es("ja,ij->ai", u1, p0["oo"]))
    return dm


def oscillator_strengths(matrix, state, level=None):
    hf = matrix.hf
    sp = {"o": "o1", "v": "v1", "c": "o2"}
    res, mus = [], []
    for vec, omega in zip(state.eigenvectors, state.eigenvalues):
        dm = transition_dm(matrix, vec, level)
        mu = np.zeros(3)
        for blk, rho in dm.items():
            s1, s2 = sp[blk[0]], sp[blk[1]]
            c1, c2 = hf.orbital_coefficients(s1), hf.orbital_coefficients(s2)
            same = hf.spin_of_axis(s1)[:, None] == hf.spin_of_axis(s2)[None, :]
            for comp in range(3):
                mo = (c1 @ hf.hf.dipole_ao[comp] @ c2.T) * same
                mu[comp] += float(np.vdot(mo, rho))
        mus.append(mu)
        res.append(2.0 / 3.0 * np.linalg.norm(mu) ** 2 * abs(omega))
    return np.array(res), np.array(mus)


def _sym(m):
    return 0.5 * (m + m.T)


def state_diffdm(matrix, vec, level=None):
    """Hermitian difference density of one state as {block: array}, blocks on and above the diagonal
    of the (o, c, v) block matrix (adcc/adc_pp/state_diffdm.py:37-160)."""
    mp, cvs = matrix.mp, matrix.cvs
    base = matrix.method[4:] if cvs else matrix.method
    if level is None:
        level = min(int(base[3]), 2)
    u1 = vec["ph"]
    u2 = vec.get("pphh")
    C = "c" if cvs else "o"
    dm = {C + C: -es("ia,ja->ij", u1, u1), "vv": es("ia,ib->ab", u1, u1)}
    if level < 1:
        return dm
    if cvs:
        if u2 is not None:
            dm["ov"] = -sqrt(2) * es("jb,ijab->ia", u1, u2)
        if level < 2:
            return dm
        t2 = mp.t2oo
        p0 = mp.mp2_diffdm(apply_cvs=True)
        p1_vv = dm["vv"]
        p2_vo = -sqrt(2) * es("ijab,jb->ai", u2, u1)
        dm["oo"] = -es("ljab,kjab->kl", u2, u2) + es("ab,ikac,jkbc->ij", p1_vv, t2, t2)
        dm["ov"] = dm["ov"] + (-es("ka,ab->kb", p0["ov"], p1_vv) - es("lkdb,dl->kb", t2, p2_vo)
                               + (1 / sqrt(2)) * es("ib,klad,liad->kb", u1, t2, u2))
        dm["vv"] = p1_vv + 2.0 * es("ijac,ijbc->ab", u2, u2) - 0.5 * (
            es("cb,ac->ab", p1_vv, p0["vv"]) + es("cb,ac->ab", p0["vv"], p1_vv)
            + es("ijbc,ijad,cd->ab", t2, t2, p1_vv))
        dm["cc"] = dm["cc"] - es("kIab,kJab->IJ", u2, u2)
        return dm
    if u2 is not None:
        dm["ov"] = -2.0 * es("jb,ijab->ia", u1, u2)
    if level < 2:
        return dm
    t2 = mp.t2oo
    p0 = mp.mp2_diffdm()
    p1_oo, p1_vv, p2_ov = dm["oo"], dm["vv"], dm["ov"]
    p2_oo = -es("ikab,jkab->ij", u2, u2)
    p2_vv = es("ijac,ijbc->ab", u2, u2)
    ru1 = es("ijab,jb->ia", t2, u1)
    dm["oo"] = p1_oo + 2.0 * p2_oo - es("ia,ja->ij", ru1, ru1) + _sym(
        es("ik,kj->ij", p1_oo, p0["oo"])
        - es("ikcd,jkcd->ij", t2, 0.5 * es("lk,jlcd->jkcd", p1_oo, t2) - es("jkcb,db->jkcd", t2, p1_vv))
        - es("ia,jkac,kc->ij", u1, t2, ru1))
    dm["vv"] = p1_vv + 2.0 * p2_vv + es("ia,ib->ab", ru1, ru1) - _sym(
        es("ac,cb->ab", p1_vv, p0["vv"])
        + es("klbc,klac->ab", t2, 0.5 * es("klad,cd->klac", t2, p1_vv) - es("jk,jlac->klac", p1_oo, t2))
        - es("ikac,kc,ib->ab", t2, ru1, u1))
    dm["ov"] = p2_ov + (-es("ijab,jb->ia", t2, p2_ov) - es("ib,ba->ia", p0["ov"], p1_vv)
                        + es("ij,ja->ia", p1_oo, p0["ov"]) - es("ib,klca,klcb->ia", u1, t2, u2)
                        - es("ikcd,jkcd,ja->ia", t2, u2, u1))
    return dm


_SP = {"o": "o1", "v": "v1", "c": "o2"}


def _dipole_block(hf, blk, comp):
    s1, s2 = _SP[blk[0]], _SP[blk[1]]
    c1, c2 = hf.orbital_coefficients(s1), hf.orbital_coefficients(s2)
    same = hf.spin_of_axis(s1)[:, None] == hf.spin_of_axis(s2)[None, :]
    return (c1 @ hf.hf.dipole_ao[comp] @ c2.T) * same


def _dipole_of_density(hf, dm):
    mu = np.zeros(3)
    for blk, rho in dm.items():
        w = 2.0 if blk[0] != blk[1] else 1.0            # Hermitian: the lower block is the transpose
        for comp in range(3):
            mu[comp] += w * float(np.vdot(_dipole_block(hf, blk, comp), rho))
    return mu


def ground_state_dipole_moment(matrix, level=2):
    """adcc/GroundState.py:137-146, 178-189."""
    hf, mp = matrix.hf, matrix.mp
    mu = np.array(hf.hf.nuclear_dipole, dtype=float)
    for sp in ["o"] + (["c"] if mp.has_core_occupied_space else []):
        for comp in range(3):
            mu[comp] += float(np.trace(_dipole_block(hf, sp + sp, comp)))
    if level >= 2:
        mu += _dipole_of_density(hf, mp.mp2_diffdm())
    return mu


def state_dipole_moments(matrix, state, level=None):
    """adcc/ElectronicStates.py:247-264."""
    cvs = matrix.cvs
    base = matrix.method[4:] if cvs else matrix.method
    if level is None:
        level = min(int(base[3]), 2)
    gs = ground_state_dipole_moment(matrix, level)
    return np.array([gs + _dipole_of_density(matrix.hf, state_diffdm(matrix, vec, level))
                     for vec in state.eigenvectors])
