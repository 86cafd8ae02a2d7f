"""Oracle transition densities / oscillator strengths (test infrastructure, see oracle/__init__.py).

Restates adcc/adc_pp/transition_dm.py:36-97 (tdm_isr0/1/2, tdm_cvs_isr2),
adcc/OperatorIntegrals.py:37-76 (AO->MO, spin-block-diagonal) and
adcc/ElectronicTransition.py:178-183 (f = 2/3 |mu|^2 omega) on dense arrays.  No in-repo known
answer exists for the offline fixture: this pins GPU-vs-oracle only.
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
