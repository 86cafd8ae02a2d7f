"""Transition densities, transition dipole moments, oscillator strengths and the
excited-state difference densities / state dipole moments that the same ExcitedStates table offers.

Mirrors adcc/adc_pp/transition_dm.py:36-140 (ISR(0), ISR(1), ISR(2), CVS-ISR(2) ground-to-excited
transition density matrices), the AO -> MO dipole transform of adcc/OperatorIntegrals.py:37-76
(block-diagonal in spin) and adcc/ElectronicTransition.py:54-80,178-183
(mu = tr(D rho), f = 2/3 |mu|^2 omega).  State difference densities follow
adcc/adc_pp/state_diffdm.py:37-160 (ISR(0), ISR(1), ISR(2), CVS-ISR(1), CVS-ISR(2)), the state dipole
moment adcc/ElectronicStates.py:247-264 and the ground-state part adcc/GroundState.py:137-146,
178-189, state-to-state transition densities adcc/adc_pp/state2state_transition_dm.py:35-118.  All
contractions run through the device einsum.
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
            u2 = _dense_doubles(amplitude)             # packed CVS store: unpack (a, b) pairs first
            t2 = mp.t2(b.oovv)
            p0 = mp.second_order_dm_correction(apply_cvs=True)
            dm["oc"] = (-1.0 * einsum("ja,Ia->jI", p0.ov, u1)
                        + (1 / sqrt(2)) * einsum("kIab,jkab->jI", u2, t2))
            dm["vc"] = dm["vc"] - 0.5 * einsum("ab,Ib->aI", p0.vv, u1)
        return dm
    if level >= 1:
        dm["ov"] = -1.0 * einsum("ijab,jb->ia", mp.t2(b.oovv), u1)
    if level >= 2:
        u2 = _dense_doubles(amplitude)
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
    if len(blk) == 4:                                  # full subspace names, e.g. "o3o3"
        s1, s2 = blk[:2], blk[2:]
    else:
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


# ------------------------------------------------------------------------------------------------
# excited-state one-particle difference densities and state dipole moments
# ------------------------------------------------------------------------------------------------
def _dense_doubles(amplitude):
    u2 = amplitude.pphh
    return u2.to_dense() if hasattr(u2, "to_dense") else u2


def state_diffdm(method, ground_state, amplitude):
    """Difference density (excited state minus ground state) of one state in the MO basis, as
    {block: Tensor}.  Hermitian: only the blocks on and above the diagonal of the (o, c, v) block
    matrix are returned ("oo", "ov", "vv", CVS: also "cc"); the lower ones are their transposes
    (adcc/adc_pp/state_diffdm.py:37-160, OneParticleDensity with OperatorSymmetry.HERMITIAN)."""
    level, cvs = _isr_level(method)
    mp = ground_state
    C = "c" if cvs else "o"
    u1 = amplitude.ph
    has_doubles = "pphh" in amplitude.blocks
    dm = {C + C: -1.0 * einsum("ia,ja->ij", u1, u1), "vv": einsum("ia,ib->ab", u1, u1)}      # ISR(0)
    if level < 1:
        return dm
    if cvs:
        if has_doubles:                                                                       # CVS-ISR(1)-d
            u2 = _dense_doubles(amplitude)
            dm["ov"] = (-sqrt(2)) * einsum("jb,ijab->ia", u1, u2)
        if level < 2:
            return dm
        if not has_doubles:
            raise ValueError("The second-order CVS difference density needs the doubles part of the vector.")
        t2 = mp.t2(b.oovv)
        p0 = mp.second_order_dm_correction(apply_cvs=True)
        p1_vv = dm["vv"]
        p2_vo = (-sqrt(2)) * einsum("ijab,jb->ai", u2, u1)
        dm["oo"] = (-1.0 * einsum("ljab,kjab->kl", u2, u2)
                    + einsum("ab,ikac,jkbc->ij", p1_vv, t2, t2))
        dm["ov"] = dm["ov"] + (-1.0 * einsum("ka,ab->kb", p0.ov, p1_vv)
                               - einsum("lkdb,dl->kb", t2, p2_vo)
                               + (1 / sqrt(2)) * einsum("ib,klad,liad->kb", u1, t2, u2))
        dm["vv"] = p1_vv + 2.0 * einsum("ijac,ijbc->ab", u2, u2) - 0.5 * (
            einsum("cb,ac->ab", p1_vv, p0.vv) + einsum("cb,ac->ab", p0.vv, p1_vv)
            + einsum("ijbc,ijad,cd->ab", t2, t2, p1_vv))
        dm["cc"] = dm["cc"] - einsum("kIab,kJab->IJ", u2, u2)
        return dm
    if has_doubles:                                                                           # ISR(1)-d
        u2 = _dense_doubles(amplitude)
        dm["ov"] = -2.0 * einsum("jb,ijab->ia", u1, u2)
    if level < 2:
        return dm
    if not has_doubles:
        raise ValueError("The second-order difference density needs the doubles part of the vector.")
    t2 = mp.t2(b.oovv)
    p0 = mp.mp2_diffdm
    p1_oo, p1_vv = dm["oo"], dm["vv"]
    p2_oo = -1.0 * einsum("ikab,jkab->ij", u2, u2)
    p2_vv = einsum("ijac,ijbc->ab", u2, u2)
    p2_ov = dm["ov"]
    ru1 = einsum("ijab,jb->ia", t2, u1)
    dm["oo"] = p1_oo + 2.0 * p2_oo - einsum("ia,ja->ij", ru1, ru1) + (
        einsum("ik,kj->ij", p1_oo, p0.oo)
        - einsum("ikcd,jkcd->ij", t2, 0.5 * einsum("lk,jlcd->jkcd", p1_oo, t2)
                 - einsum("jkcb,db->jkcd", t2, p1_vv))
        - einsum("ia,jkac,kc->ij", u1, t2, ru1)).symmetrise()
    dm["vv"] = p1_vv + 2.0 * p2_vv + einsum("ia,ib->ab", ru1, ru1) - (
        einsum("ac,cb->ab", p1_vv, p0.vv)
        + einsum("klbc,klac->ab", t2, 0.5 * einsum("klad,cd->klac", t2, p1_vv)
                 - einsum("jk,jlac->klac", p1_oo, t2))
        - einsum("ikac,kc,ib->ab", t2, ru1, u1)).symmetrise()
    dm["ov"] = p2_ov + (-1.0 * einsum("ijab,jb->ia", t2, p2_ov)
                        - einsum("ib,ba->ia", p0.ov, p1_vv)
                        + einsum("ij,ja->ia", p1_oo, p0.ov)
                        - einsum("ib,klca,klcb->ia", u1, t2, u2)
                        - einsum("ikcd,jkcd,ja->ia", t2, u2, u1))
    return dm


def _dipole_of_density(refstate, dm, cache=None, hermitian=True):
    """tr(D rho) per Cartesian component for a block density; ``hermitian``: only the blocks on and
    above the diagonal are stored, each off-diagonal one counts twice (D is symmetric)."""
    cache = {} if cache is None else cache
    mu = np.zeros(3)
    for comp in range(3):
        for blk, rho in dm.items():
            if rho is None:
                continue
            key = (blk, comp)
            if key not in cache:
                cache[key] = dipole_mo_block(refstate, blk, comp)
            rho._contig()
            mu[comp] += (2.0 if hermitian and blk[0] != blk[1] else 1.0) * cache[key].dot(rho)
    return mu


def ground_state_dipole_moment(ground_state, level=2, cache=None):
    """Nuclear + electronic dipole moment of the MP ground state with the density corrected up to
    ``level`` (adcc/GroundState.py:137-146, 178-189): the HF density is the identity on the occupied
    (and core) block, level >= 2 adds the second-order correction (all blocks, no CVS truncation)."""
    ref = ground_state.reference_state
    cache = {} if cache is None else cache
    ms = ref.mospaces
    mu = np.array(ref.nuclear_dipole, dtype=float)
    for sp in ms.subspaces_occupied:                    # o1, o2 (core), o3 (frozen): the HF density is the identity there
        for comp in range(3):
            key = (sp + sp, comp)
            if key not in cache:
                cache[key] = dipole_mo_block(ref, sp + sp, comp)
            mu[comp] += float(np.trace(cache[key].to_ndarray()))
    if level >= 2:
        corr = ground_state.diffdm(2)
        dm = {blk: corr[blk] for blk in ("oo", "ov", "vv", "cc", "oc", "cv") if blk in corr}
        mu += _dipole_of_density(ref, dm, cache)
    return mu


def state_dipole_moments(states):
    """Dipole moments of the excited states (adcc/ElectronicStates.py:247-264): ground-state dipole
    at the level of the property method plus tr(D diffdm) of each state."""
    level, _ = _isr_level(states.property_method)
    cache = {}
    gs = ground_state_dipole_moment(states.ground_state, level, cache)
    return np.array([gs + _dipole_of_density(states.reference_state,
                                             state_diffdm(states.property_method, states.ground_state, v), cache)
                     for v in states.excitation_vector])


# ------------------------------------------------------------------------------------------------
# transition densities between excited states
# ------------------------------------------------------------------------------------------------
def state2state_transition_dm(method, ground_state, amplitude_from, amplitude_to):
    """Transition density between two excited states as {block: Tensor} with the blocks "oo", "vv"
    and - from first order with doubles - "ov", "vo" (no symmetry between them).  The final state is
    the bra side (adcc/adc_pp/state2state_transition_dm.py:35-118, 161-165)."""
    level, cvs = _isr_level(method)
    if cvs:
        raise NotImplementedError("state2state_transition_dm is not implemented for CVS methods.")
    mp = ground_state
    left, right = amplitude_to, amplitude_from
    ul1, ur1 = left.ph, right.ph
    dm = {"oo": -1.0 * einsum("ja,ia->ij", ul1, ur1), "vv": einsum("ia,ib->ab", ul1, ur1)}
    has_doubles = "pphh" in left.blocks and "pphh" in right.blocks
    if level < 1:
        return dm
    if has_doubles:
        ul2, ur2 = _dense_doubles(left), _dense_doubles(right)
        dm["ov"] = -2.0 * einsum("jb,ijab->ia", ul1, ur2)
        dm["vo"] = -2.0 * einsum("ijab,jb->ai", ul2, ur1)
    if level < 2:
        return dm
    if not has_doubles:
        raise ValueError("The second-order state-to-state density needs the doubles parts of both vectors.")
    t2 = mp.t2(b.oovv)
    p0 = mp.mp2_diffdm
    p1_oo, p1_vv, p1_ov, p1_vo = dm["oo"], dm["vv"], dm["ov"], dm["vo"]
    rul1 = einsum("ijab,jb->ia", t2, ul1)
    rur1 = einsum("ijab,jb->ia", t2, ur1)
    dm["oo"] = p1_oo + (
        -2.0 * einsum("ikab,jkab->ij", ur2, ul2)
        + 0.5 * einsum("ik,kj->ij", p1_oo, p0.oo) + 0.5 * einsum("ik,kj->ij", p0.oo, p1_oo)
        - 0.5 * einsum("ikcd,lk,jlcd->ij", t2, p1_oo, t2)
        + 1.0 * einsum("ikcd,jkcb,db->ij", t2, t2, p1_vv)
        - 0.5 * einsum("ia,jkac,kc->ij", ur1, t2, rul1)
        - 0.5 * einsum("ikac,kc,ja->ij", t2, rur1, ul1)
        - 1.0 * einsum("ia,ja->ij", rul1, rur1))
    dm["vv"] = p1_vv + (
        2.0 * einsum("ijac,ijbc->ab", ul2, ur2)
        - 0.5 * einsum("ac,cb->ab", p1_vv, p0.vv) - 0.5 * einsum("ac,cb->ab", p0.vv, p1_vv)
        - 0.5 * einsum("klbc,klad,cd->ab", t2, t2, p1_vv)
        + 1.0 * einsum("klbc,jk,jlac->ab", t2, p1_oo, t2)
        + 0.5 * einsum("ikac,kc,ib->ab", t2, rul1, ur1)
        + 0.5 * einsum("ia,ikbc,kc->ab", ul1, t2, rur1)
        + 1.0 * einsum("ia,ib->ab", rur1, rul1))
    dm["ov"] = p1_ov + (
        -1.0 * einsum("ijab,bj->ia", t2, p1_vo)
        - einsum("ib,ba->ia", p0.ov, p1_vv)
        + einsum("ij,ja->ia", p1_oo, p0.ov)
        - einsum("ib,klca,klcb->ia", ur1, t2, ul2)
        - einsum("ikcd,jkcd,ja->ia", t2, ul2, ur1))
    dm["vo"] = p1_vo + (
        -1.0 * einsum("ijab,jb->ai", t2, p1_ov)
        - einsum("ib,ab->ai", p0.ov, p1_vv)
        + einsum("ji,ja->ai", p1_oo, p0.ov)
        - einsum("ib,klca,klcb->ai", ul1, t2, ur2)
        - einsum("ikcd,jkcd,ja->ai", t2, ur2, ul1))
    return dm
