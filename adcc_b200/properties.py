"""Transition densities, transition dipole moments and oscillator strengths.

Mirrors adcc/adc_pp/transition_dm.py:36-140 (ISR(0), ISR(1), ISR(2), CVS-ISR(2) ground-to-excited
transition density matrices), the AO -> MO dipole transform of adcc/OperatorIntegrals.py:37-76
(block-diagonal in spin) and adcc/ElectronicTransition.py:54-80,178-183
(mu = tr(D rho), f = 2/3 |mu|^2 omega).  All contractions run through the device einsum.
"""
from math import sqrt

import numpy as np

from . import block as b
from .AdcMethod import AdcMethod
from .functions import einsum
from .Tensor import Tensor


def _isr_level(method):
    m = AdcMethod(method) if not isinstance(method, AdcMethod) else method
    return min(m.level, 2), m.is_core_valence_separated


def transition_dm(method, ground_state, amplitude):
    """dict {block: Tensor} with blocks named like adcc's OneParticleDensity ("vo", "ov", ...)."""
    level, cvs = _isr_level(method)
    mp = ground_state
    C = "c" if cvs else "o"
    u1 = amplitude.ph
    dm = {"v" + C: u1.transpose()}
    if cvs:
        if level >= 2:
            u2 = amplitude.pphh
            t2 = mp.t2(b.oovv)
            p0 = mp.second_order_dm_correction(apply_cvs=True)
            dm["oc"] = (-1.0 * einsum("ja,Ia->jI", p0.ov, u1)
                        + (1 / sqrt(2)) * einsum("kIab,jkab->jI", u2, t2))
            dm["vc"] = dm["vc"] - 0.5 * einsum("ab,Ib->aI", p0.vv, u1)
        return dm
    if level >= 1:
        dm["ov"] = -1.0 * einsum("ijab,jb->ia", mp.t2(b.oovv), u1)
    if level >= 2:
        u2 = amplitude.pphh
        if hasattr(u2, "to_dense"):
            u2 = u2.to_dense()
        t2 = mp.t2(b.oovv)
        td2 = mp.td2(b.oovv)
        p0 = mp.mp2_diffdm
        dm["oo"] = (-1.0 * einsum("ia,ja->ij", p0.ov, u1) - einsum("ikab,jkab->ji", u2, t2))
        dm["vv"] = (einsum("ia,ib->ab", u1, p0.ov) + einsum("ijac,ijbc->ab", u2, t2))
        dm["ov"] = dm["ov"] - einsum("ijab,jb->ia", td2, u1)
        dm["vo"] = dm["vo"] + 0.5 * (einsum("ijab,jkbc,kc->ai", t2, t2, u1)
                                     - einsum("ab,ib->ai", p0.vv, u1)
                                     + einsum("ja,ij->ai", u1, p0.oo))
    return dm


_SPACE = {"o": "o1", "v": "v1", "c": "o2"}


def dipole_mo_block(refstate, blk, component):
    """<p| r_component |q> for p in space blk[0], q in space blk[1]; cross-spin elements vanish."""
    prov = refstate.operator_integral_provider
    if prov is None or not hasattr(prov, "electric_dipole"):
        raise NotImplementedError("The SCF data provides no electric dipole integrals")
    s1, s2 = _SPACE[blk[0]], _SPACE[blk[1]]
    d_ao = np.asarray(prov.electric_dipole[component])
    c1 = refstate.orbital_coefficients_host(s1)
    c2 = refstate.orbital_coefficients_host(s2)
    mo = c1 @ d_ao @ c2.T
    ms = refstate.mospaces
    spin1 = np.arange(ms.n_orbs(s1)) >= ms.n_orbs_alpha(s1)
    spin2 = np.arange(ms.n_orbs(s2)) >= ms.n_orbs_alpha(s2)
    mo = mo * (spin1[:, None] == spin2[None, :])
    return Tensor.from_ndarray(ms, [s1, s2], mo)


def transition_dipole_moments(states):
    ref = states.reference_state
    out = []
    cache = {}
    for vec in states.excitation_vector:
        dm = transition_dm(states.property_method, states.ground_state, vec)
        mu = np.zeros(3)
        for comp in range(3):
            for blk, rho in dm.items():
                key = (blk, comp)
                if key not in cache:
                    cache[key] = dipole_mo_block(ref, blk, comp)
                rho._contig()
                mu[comp] += cache[key].dot(rho)
        out.append(mu)
    return np.array(out)
