"""Oracle transition densities / oscillator strengths / state densities (test infrastructure, see
oracle/__init__.py).

Restates adcc/adc_pp/transition_dm.py:36-97 (tdm_isr0/1/2, tdm_cvs_isr2),
adcc/OperatorIntegrals.py:37-76 (AO->MO, spin-block-diagonal) and
adcc/ElectronicTransition.py:178-183 (f = 2/3 |mu|^2 omega) on dense arrays.  State difference densities and
state dipole moments restate adcc/adc_pp/state_diffdm.py:37-160, adcc/ElectronicStates.py:247-264 and
adcc/GroundState.py:137-146, 178-189; state-to-state transition densities restate
adcc/adc_pp/state2state_transition_dm.py:35-118 and adcc/State2States.py:79-140.  Pinned by the goldens the reference's own property code
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
    s1, s2 = (blk[:2], blk[2:]) if len(blk) == 4 else (_SP[blk[0]], _SP[blk[1]])
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
    for sp in ("o1", "o2", "o3"):                    # active, core and frozen occupied: HF density = identity
        if hf.n_orbs.get(sp, 0) == 0:
            continue
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


def state2state_transition_dm(matrix, vec_from, vec_to, level=None):
    """adcc/adc_pp/state2state_transition_dm.py:35-118 (the final state is the bra side, :161-165)."""
    mp = matrix.mp
    if matrix.cvs:
        raise NotImplementedError("no CVS state-to-state densities in the reference")
    if level is None:
        level = min(int(matrix.method[3]), 2)
    ul1, ur1 = vec_to["ph"], vec_from["ph"]
    ul2, ur2 = vec_to.get("pphh"), vec_from.get("pphh")
    dm = {"oo": -es("ja,ia->ij", ul1, ur1), "vv": es("ia,ib->ab", ul1, ur1)}
    if level < 1:
        return dm
    if ul2 is not None and ur2 is not None:
        dm["ov"] = -2.0 * es("jb,ijab->ia", ul1, ur2)
        dm["vo"] = -2.0 * es("ijab,jb->ai", ul2, ur1)
    if level < 2:
        return dm
    t2 = mp.t2oo
    p0 = mp.mp2_diffdm()
    p1_oo, p1_vv, p1_ov, p1_vo = dm["oo"], dm["vv"], dm["ov"], dm["vo"]
    rul1 = es("ijab,jb->ia", t2, ul1)
    rur1 = es("ijab,jb->ia", t2, ur1)
    dm["oo"] = p1_oo + (
        -2.0 * es("ikab,jkab->ij", ur2, ul2)
        + 0.5 * es("ik,kj->ij", p1_oo, p0["oo"]) + 0.5 * es("ik,kj->ij", p0["oo"], p1_oo)
        - 0.5 * es("ikcd,lk,jlcd->ij", t2, p1_oo, t2) + es("ikcd,jkcb,db->ij", t2, t2, p1_vv)
        - 0.5 * es("ia,jkac,kc->ij", ur1, t2, rul1) - 0.5 * es("ikac,kc,ja->ij", t2, rur1, ul1)
        - es("ia,ja->ij", rul1, rur1))
    dm["vv"] = p1_vv + (
        2.0 * es("ijac,ijbc->ab", ul2, ur2)
        - 0.5 * es("ac,cb->ab", p1_vv, p0["vv"]) - 0.5 * es("ac,cb->ab", p0["vv"], p1_vv)
        - 0.5 * es("klbc,klad,cd->ab", t2, t2, p1_vv) + es("klbc,jk,jlac->ab", t2, p1_oo, t2)
        + 0.5 * es("ikac,kc,ib->ab", t2, rul1, ur1) + 0.5 * es("ia,ikbc,kc->ab", ul1, t2, rur1)
        + es("ia,ib->ab", rur1, rul1))
    dm["ov"] = p1_ov + (
        -es("ijab,bj->ia", t2, p1_vo) - es("ib,ba->ia", p0["ov"], p1_vv) + es("ij,ja->ia", p1_oo, p0["ov"])
        - es("ib,klca,klcb->ia", ur1, t2, ul2) - es("ikcd,jkcd,ja->ia", t2, ul2, ur1))
    dm["vo"] = p1_vo + (
        -es("ijab,jb->ai", t2, p1_ov) - es("ib,ab->ai", p0["ov"], p1_vv) + es("ji,ja->ai", p1_oo, p0["ov"])
        - es("ib,klca,klcb->ai", ul1, t2, ur2) - es("ikcd,jkcd,ja->ai", t2, ur2, ul1))
    return dm


def state2state(matrix, state, initial=0, level=None):
    """(excitation energies, transition dipole moments, oscillator strengths, densities) from state
    ``initial`` to every higher one (adcc/State2States.py:79-140)."""
    hf = matrix.hf
    e0, v0 = state.eigenvalues[initial], state.eigenvectors[initial]
    ene, mus, dms = [], [], []
    for e, v in zip(state.eigenvalues[initial + 1:], state.eigenvectors[initial + 1:]):
        dm = state2state_transition_dm(matrix, v0, v, level)
        mu = np.zeros(3)
        for blk, rho in dm.items():
            for comp in range(3):
                mu[comp] += float(np.vdot(_dipole_block(hf, blk, comp), rho))
        ene.append(e - e0)
        mus.append(mu)
        dms.append(dm)
    ene, mus = np.array(ene), np.array(mus).reshape(len(ene), 3)
    f = 2.0 / 3.0 * np.array([np.linalg.norm(m) ** 2 * abs(w) for m, w in zip(mus, ene)])
    return ene, mus, f, dms
