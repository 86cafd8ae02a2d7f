"""Shared parity checks: product path (whatever backend is active) versus the oracle."""
import numpy as np

import oracle

METHODS = ["adc0", "adc1", "adc2", "adc2x", "adc3",
           "cvs-adc0", "cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3"]
SINGLETS = {
    "adc0": [0.99048446, 1.04450741, 1.15356339], "adc1": [0.48343609, 0.57420044, 0.60213699],
    "adc2": [0.47051314, 0.57255495, 0.59367339], "adc2x": [0.44511182, 0.55139508, 0.57146893],
    "adc3": [0.45197068, 0.55572706, 0.57861241],
    "cvs-adc0": [20.83623681, 20.99931574], "cvs-adc1": [20.10577515, 20.16512502],
    "cvs-adc2": [20.0045422, 20.08771799], "cvs-adc2x": [19.90528006, 19.9990468],
    "cvs-adc3": [19.93138526, 20.02055267],
}
SIGMA_RTOL = 1e-11      # north star: sigma vectors within 1e-11 relative
ENERGY_ATOL = 1e-8      # north star: excitation energies within 1e-8 Hartree


def build_pair(data, method, core):
    import adcc_b200 as adcc
    ref = adcc.ReferenceState(data, core_orbitals=core)
    matrix = adcc.AdcMatrix(method, ref)
    oref = oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core)
    omatrix = oracle.OracleAdcMatrix(method, oref)
    return matrix, omatrix


def random_vector(matrix, omatrix, seed):
    import adcc_b200 as adcc
    rng = np.random.default_rng(seed)
    ov = {k: rng.standard_normal(s) for k, s in omatrix.shapes.items()}
    if "pphh" in ov:
        ov["pphh"] = omatrix.symmetrise_doubles(ov["pphh"])
    v = adcc.AmplitudeVector(**{
        k: adcc.Tensor.from_ndarray(matrix.mospaces, matrix.axis_spaces[k], arr) for k, arr in ov.items()})
    return v, ov


def rel_err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def check_matvec_parity(data, method, core, seed=42):
    matrix, omatrix = build_pair(data, method, core)
    return compare_matrices(matrix, omatrix, seed)


def compare_matrices(matrix, omatrix, seed=42):
    """diagonal, sigma, per-block applies and hermiticity of a dense-store product matrix against an
    oracle matrix of the same method / reference (adcc/tests/AdcMatrix_test.py:213-285)."""
    # diagonal
    for k, d in omatrix.diagonal().items():
        assert rel_err(matrix.diagonal()[k].to_ndarray(), d) < SIGMA_RTOL, f"diagonal {k}"
    v, ov = random_vector(matrix, omatrix, seed)
    sig = matrix.matvec(v)
    osig = omatrix.matvec(ov)
    assert sorted(sig.blocks) == sorted(osig)
    for k in osig:
        assert rel_err(sig[k].to_ndarray(), osig[k]) < SIGMA_RTOL, f"sigma {k}"
    # per-block applies (AdcMatrix.block_apply, adcc/tests/AdcMatrix_test.py:256-285)
    for blk in omatrix.blocks:
        outb, inb = blk.split("_")
        if inb not in ov or matrix.block_orders[blk] in (None, 0) and outb != inb:
            continue
        ores = omatrix.blocks[blk]({inb: ov[inb]})
        if ores is None:
            continue
        res = matrix.block_apply(blk, v[inb])
        assert rel_err(res.to_ndarray(), ores[outb]) < SIGMA_RTOL, f"block {blk}"
    # hermiticity <u|Mw> = <Mu|w>
    w, _ = random_vector(matrix, omatrix, seed + 1)
    lhs, rhs = v @ matrix.matvec(w), sig @ w
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs))
    return matrix, omatrix


def check_run_parity(data, method, kind, core, n=None, conv_tol=1e-9):
    import adcc_b200 as adcc
    if n is None:
        n = len(SINGLETS[method]) if method in SINGLETS else 3
    kw = {"singlet": {"n_singlets": n}, "triplet": {"n_triplets": n}, "any": {"n_states": n}}[kind]
    state = adcc.run_adc(data, method=method, conv_tol=conv_tol, core_orbitals=core, output=None, **kw)
    okw = dict(kw)
    ostate = oracle.run_adc(data, method, conv_tol=conv_tol, core_orbitals=core,
                            kind="any" if kind == "any" else kind, **okw)
    assert state.converged and ostate.converged
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    assert state.n_iter == ostate.n_iter
    return state, ostate


def check_oscillator_strengths(data, method, core, n=3, conv_tol=1e-9):
    """transition dipoles / oscillator strengths, product vs oracle (<= 1e-6 rel)."""
    import adcc_b200 as adcc
    from oracle.properties import oscillator_strengths
    n = 2 if method.startswith("cvs") else n
    state = adcc.run_adc(data, method=method, n_singlets=n, conv_tol=conv_tol, core_orbitals=core, output=None)
    ostate = oracle.run_adc(data, method, n_singlets=n, conv_tol=conv_tol, core_orbitals=core)
    assert state.converged and ostate.converged
    oref = oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core)
    of, omu = oscillator_strengths(oracle.OracleAdcMatrix(method, oref), ostate)
    f = state.oscillator_strength
    assert f.shape == of.shape
    np.testing.assert_allclose(f, of, rtol=1e-6, atol=1e-10)
    assert np.all(f >= 0)
    # state difference densities / state dipole moments (adcc/adc_pp/state_diffdm.py, ElectronicStates.py:247-264)
    from oracle.properties import state_diffdm, state_dipole_moments
    om = oracle.OracleAdcMatrix(method, oref)
    tol = max(1e-7, 100.0 * conv_tol)                 # first order in the eigenvector error
    np.testing.assert_allclose(state.state_dipole_moment, state_dipole_moments(om, ostate), rtol=0, atol=tol)
    for dm, ovec in zip(state.state_diffdm, ostate.eigenvectors):
        odm = state_diffdm(om, ovec)
        assert sorted(dm) == sorted(odm)
        for blk in dm:
            assert np.max(np.abs(dm[blk].to_ndarray() - odm[blk])) < tol, (method, blk)
        # a difference density moves no charge: tr(diffdm) = 0 (second order: to the order of the method)
        tr = sum(float(np.trace(dm[blk].to_ndarray())) for blk in dm if blk[0] == blk[1])
        assert abs(tr) < (1e-10 if method in ("adc0", "adc1", "cvs-adc0", "cvs-adc1") else 0.2)
    if not method.startswith("cvs") and state.size > 1:
        # transitions between excited states (adcc/State2States.py, adc_pp/state2state_transition_dm.py)
        from oracle.properties import state2state
        s2s = adcc.State2States(state, initial=0)
        oene, omus, of2, _ = state2state(om, ostate, initial=0)
        np.testing.assert_allclose(s2s.excitation_energy, oene, rtol=0, atol=1e-8)
        np.testing.assert_allclose(np.abs(s2s.transition_dipole_moment), np.abs(omus), rtol=0, atol=tol)
        np.testing.assert_allclose(s2s.oscillator_strength, of2, rtol=1e-4, atol=10 * tol)
        assert s2s.size == state.size - 1 and len(s2s.transition_dm) == s2s.size
    text = state.describe(state_dipole_moments=True)
    assert "state dipole moment" in text
    return f


def check_shifted_projected(data):
    import adcc_b200 as adcc
    matrix, omatrix = build_pair(data, "adc2x", None)
    v, ov = random_vector(matrix, omatrix, 7)
    osig = omatrix.matvec(ov)
    # shifted
    sh = adcc.AdcMatrixShifted(matrix, shift=-0.35)
    ssig = sh.matvec(v)
    for k in osig:
        assert rel_err(ssig[k].to_ndarray(), osig[k] - 0.35 * ov[k]) < SIGMA_RTOL
        assert rel_err(sh.diagonal()[k].to_ndarray(), omatrix.diagonal()[k] - 0.35) < SIGMA_RTOL
    # projected onto core excitations (1 core orbital per spin)
    pm = adcc.AdcMatrixProjected(matrix, ["cv", "ccvv", "ocvv"], core_orbitals=1)
    no, na = omatrix.shapes["ph"][0], omatrix.hf.n_alpha_of["o1"]
    core = np.zeros(no)
    core[[0, na]] = 1.0
    val = 1.0 - core
    m1 = core[:, None] * np.ones(omatrix.shapes["ph"][1])[None, :]
    m2o = np.maximum(np.maximum(np.multiply.outer(core, core), np.multiply.outer(val, core)),
                     np.multiply.outer(core, val))
    m2 = m2o[:, :, None, None] * np.ones(omatrix.shapes["pphh"][2:])[None, None]
    masks = {"ph": m1, "pphh": m2}
    for k in masks:
        assert np.array_equal(pm.projectors[k].kernel.to_ndarray(), masks[k])
    pov = {k: masks[k] * ov[k] for k in ov}
    posig = omatrix.matvec(pov)
    psig = pm.matvec(v)
    for k in posig:
        assert rel_err(psig[k].to_ndarray(), masks[k] * posig[k]) < SIGMA_RTOL
        expect_d = masks[k] * omatrix.diagonal()[k] + 100000 * (1 - masks[k])
        assert rel_err(pm.diagonal()[k].to_ndarray(), expect_d) < SIGMA_RTOL
    # core-excited states through the projected full matrix: lowest singlet close to CVS-ADC(2)-x
    state = adcc.run_adc(pm, n_singlets=2, n_guesses=2, conv_tol=1e-7, output=None)
    assert state.converged
    assert abs(state.excitation_energy[0] - 19.90528006) < 0.05
    return state


def ao_version_of(data, seed=3):
    """Same SCF data expressed in a rotated 'AO' basis: AO ERIs + non-trivial MO coefficients."""
    from oracle.hfdata import OracleHf
    na = np.asarray(data["orbcoeff_fb"]).shape[0] // 2
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((na, na)))     # orthogonal: C = q, AO = back-rotated MO
    if "eri_ffff_spatial" in data:
        mo = np.asarray(data["eri_ffff_spatial"])
    else:
        mo = np.asarray(data["eri_ffff"])[:na, :na, :na, :na]
    ao = np.einsum("pqrs,pm,qn,rl,sk->mnlk", mo, q, q, q, q, optimize=True)
    out = {k: v for k, v in data.items() if k not in ("eri_ffff", "eri_ffff_spatial")}
    out["eri_bbbb"] = ao
    out["orbcoeff_fb"] = np.vstack([q, q])
    if "multipoles" in data and "elec_1" in data["multipoles"]:
        mm = dict(data["multipoles"])
        c_old = np.asarray(data["orbcoeff_fb"])[:na]
        d_mo = np.einsum("pm,xmn,qn->xpq", c_old, mm["elec_1"], c_old)
        mm["elec_1"] = np.einsum("xpq,pm,qn->xmn", d_mo, q, q)
        out["multipoles"] = mm
    return out


def check_ao_import(data, method="adc2", core=None):
    """8(f4): ERIs given in the AO basis, transformed on the device, must give the same matrix."""
    import adcc_b200 as adcc
    ao = ao_version_of(data)
    ref_ao = adcc.ReferenceState(ao, core_orbitals=core)
    ref_mo = adcc.ReferenceState(data, core_orbitals=core)
    for blk in ("o1o1v1v1", "o1v1o1v1", "o1v1v1v1", "v1v1v1v1", "o1o1o1v1"):
        a, b = ref_ao.eri(blk).to_ndarray(), ref_mo.eri(blk).to_ndarray()
        assert np.max(np.abs(a - b)) < 1e-12 * max(1.0, np.max(np.abs(b))), blk
    state = adcc.run_adc(ref_ao, method=method, n_singlets=2, conv_tol=1e-8, output=None)
    ostate = oracle.run_adc(data, method, n_singlets=2, conv_tol=1e-8, core_orbitals=core)
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    return state


def check_unrestricted_kinds(no=6, nv=8):
    import adcc_b200 as adcc
    from adcc_b200.packed import synthetic_reference_state
    from oracle.synthetic import OracleSyntheticRef
    ref = synthetic_reference_state(no, nv, seed=9, scale=0.02)
    assert not ref.restricted
    with pytest_raises(adcc.InputError):
        adcc.run_adc(ref, method="adc2", n_singlets=2, output=None)       # restricted-only keyword
    for method in ("adc1", "adc2"):
        om = oracle.OracleAdcMatrix(method, OracleSyntheticRef(no, nv, seed=9, scale=0.02))
        m = adcc.AdcMatrix(method, ref, doubles_storage="dense")
        st = adcc.run_adc(m, n_states=3, conv_tol=1e-8, output=None)
        ost = oracle.run_adc(om, method, n_states=3, conv_tol=1e-8)
        np.testing.assert_allclose(st.excitation_energy, ost.eigenvalues, atol=ENERGY_ATOL, rtol=0)
        sf = adcc.run_adc(m, n_spin_flip=2, conv_tol=1e-8, output=None)
        osf = oracle.run_adc(om, method, n_states=2, kind="spin_flip", conv_tol=1e-8)
        assert sf.converged and osf.converged
        np.testing.assert_allclose(sf.excitation_energy, osf.eigenvalues, atol=ENERGY_ATOL, rtol=0)


class pytest_raises:
    def __init__(self, exc): self.exc = exc
    def __enter__(self): return self
    def __exit__(self, et, ev, tb):
        if et is None or not issubclass(et, self.exc):
            raise AssertionError(f"{self.exc} not raised")
        return True


def check_batched_matvec(data, method, core):
    """8(f2): matrix @ [v1..vn] (stacked, one GEMM per term) equals the per-vector products."""
    import adcc_b200 as adcc
    matrix, omatrix = build_pair(data, method, core)
    vecs = [random_vector(matrix, omatrix, 100 + k) for k in range(3)]
    batched = matrix @ [v for v, _ in vecs]
    assert len(batched) == 3
    for (v, ov), res in zip(vecs, batched):
        osig = omatrix.matvec(ov)
        single = matrix.matvec(v)
        for k in osig:
            assert rel_err(res[k].to_ndarray(), osig[k]) < SIGMA_RTOL
            assert rel_err(res[k].to_ndarray(), single[k].to_ndarray()) < 1e-13


def open_shell_dict(n_orb=6, n_alpha=3, n_beta=2, seed=4, scale=0.03):
    """Unrestricted open-shell reference (n_alpha != n_beta) given directly as <pq||rs>
    (key eri_phys_asym_ffff, adcc/DataHfProvider.py:121-132)."""
    rng = np.random.default_rng(seed)
    nf = 2 * n_orb
    g = rng.standard_normal((nf,) * 4) * scale
    g = g - g.transpose(1, 0, 2, 3)
    g = g - g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    ea = np.sort(rng.uniform(-1.5, -0.4, n_alpha)).tolist() + np.sort(rng.uniform(0.2, 2.5, n_orb - n_alpha)).tolist()
    eb = np.sort(rng.uniform(-1.5, -0.4, n_beta)).tolist() + np.sort(rng.uniform(0.2, 2.5, n_orb - n_beta)).tolist()
    orben = np.array(ea + eb)
    occ = np.array([1.0] * n_alpha + [0.0] * (n_orb - n_alpha) + [1.0] * n_beta + [0.0] * (n_orb - n_beta))
    return {"restricted": False, "conv_tol": 1e-12, "spin_multiplicity": n_alpha - n_beta + 1,
            "orbcoeff_fb": np.vstack([np.eye(n_orb)] * 2), "occupation_f": occ, "orben_f": orben,
            "fock_ff": np.diag(orben), "eri_phys_asym_ffff": g}


def synthetic_workload_dict(no, nv, seed, scale):
    """The synthetic workload of BASELINE configs[4] (counter-based <pq||rs>, adcc_b200/synthetic.py) as
    a plain SCF dict with the complete antisymmetrised array ``eri_phys_asym_ffff`` over all
    no + nv spin orbitals, for consumers that know nothing about the packed source (the reference's
    DataHfProvider).  The six defined blocks are completed by <pq||rs> = -<qp||rs> = -<pq||sr> = <rs||pq>;
    host order of the orbital axis: alpha (occupied, virtual), then beta (adcc/DataHfProvider.py:84-141)."""
    from adcc_b200.synthetic import synthetic_eri_block, synthetic_orbital_energies
    nf = no + nv
    O, V = slice(0, no), slice(no, nf)
    B = {k: synthetic_eri_block(seed, k, no, nv, scale) for k in ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")}
    G = np.zeros((nf,) * 4)
    G[O, O, O, O] = B["oooo"]
    G[O, O, O, V] = B["ooov"]
    G[O, O, V, O] = -B["ooov"].transpose(0, 1, 3, 2)
    G[O, V, O, O] = B["ooov"].transpose(2, 3, 0, 1)
    G[V, O, O, O] = -B["ooov"].transpose(3, 2, 0, 1)
    G[O, O, V, V] = B["oovv"]
    G[V, V, O, O] = B["oovv"].transpose(2, 3, 0, 1)
    G[O, V, O, V] = B["ovov"]
    G[O, V, V, O] = -B["ovov"].transpose(0, 1, 3, 2)
    G[V, O, O, V] = -B["ovov"].transpose(1, 0, 2, 3)
    G[V, O, V, O] = B["ovov"].transpose(1, 0, 3, 2)
    G[O, V, V, V] = B["ovvv"]
    G[V, O, V, V] = -B["ovvv"].transpose(1, 0, 2, 3)
    G[V, V, O, V] = B["ovvv"].transpose(2, 3, 0, 1)
    G[V, V, V, O] = -B["ovvv"].transpose(2, 3, 1, 0)
    G[V, V, V, V] = B["vvvv"]
    assert np.max(np.abs(G + G.transpose(1, 0, 2, 3))) == 0 and np.max(np.abs(G + G.transpose(0, 1, 3, 2))) == 0
    assert np.max(np.abs(G - G.transpose(2, 3, 0, 1))) == 0
    h2i = np.concatenate((np.arange(no // 2), no + np.arange(nv // 2),
                          no // 2 + np.arange(no // 2), no + nv // 2 + np.arange(nv // 2)))
    e_o, e_v = synthetic_orbital_energies(no, nv)
    orben = np.concatenate((e_o[0::2], e_v[0::2], e_o[1::2], e_v[1::2]))     # as adcc_b200.packed.synthetic_reference_state
    na = nf // 2
    occ = np.concatenate((np.ones(no // 2), np.zeros(nv // 2), np.ones(no // 2), np.zeros(nv // 2)))
    return {"restricted": False, "conv_tol": 1e-12, "energy_scf": 0.0, "spin_multiplicity": 1,
            "orbcoeff_fb": np.vstack([np.eye(na)] * 2), "occupation_f": occ, "orben_f": orben,
            "fock_ff": np.diag(orben), "eri_phys_asym_ffff": G[np.ix_(h2i, h2i, h2i, h2i)]}


def check_open_shell():
    data = open_shell_dict()
    for method in ("adc2", "adc2x", "adc3"):
        matrix, omatrix = check_matvec_parity(data, method, None)
        assert matrix.mospaces.n_orbs_alpha("o1") == 3 and matrix.mospaces.n_orbs_beta("o1") == 2
    check_run_parity(data, "adc2", "any", None, n=3, conv_tol=1e-8)


def check_graph_replay(data, method, core, monkeypatch):
    """Captured-graph matvec (single and batched) equals the eagerly launched one on fresh inputs
    each replay, and the oracle."""
    matrix, omatrix = build_pair(data, method, core)
    matrix.GRAPH_CAPTURE_AFTER = 1
    vecs = [random_vector(matrix, omatrix, 300 + k) for k in range(5)]
    monkeypatch.setenv("ADCCB_GRAPHS", "0")
    eager = [matrix.matvec(v) for v, _ in vecs]
    assert not matrix._graphs
    monkeypatch.setenv("ADCCB_GRAPHS", "1")
    replayed = [matrix.matvec(v) for v, _ in vecs]                # eager, capture, 3 replays
    sets = [vecs[:3], vecs[3:], vecs[1:4], vecs[:2], vecs[2:5]]   # n=3: eager, capture, replay
    batched = [matrix @ [v for v, _ in vs] for vs in sets]
    assert set(matrix._graphs) == {(n, tuple(vecs[0][0].blocks)) for n in (1, 2, 3)}
    for k, (v, ov) in enumerate(vecs):
        osig = omatrix.matvec(ov)
        for b in osig:
            assert rel_err(replayed[k][b].to_ndarray(), osig[b]) < SIGMA_RTOL
            assert np.array_equal(replayed[k][b].to_ndarray(), eager[k][b].to_ndarray())
    for vs, res in zip(sets, batched):
        for (v, ov), sig in zip(vs, res):
            k = [id(x) for x, _ in vecs].index(id(v))
            for b in eager[k].blocks:
                assert rel_err(sig[b].to_ndarray(), eager[k][b].to_ndarray()) < 1e-13
    # outputs are copies: a later replay must not overwrite an earlier result
    first = replayed[2]["ph"].to_ndarray().copy()
    matrix.matvec(vecs[1][0])
    assert np.array_equal(replayed[2]["ph"].to_ndarray(), first)


def check_frozen_spaces(data, method, core, frozen_core, frozen_virtual, n=2):
    """Frozen-core / frozen-virtual subspaces (the reference's fc, fv, fc-fv, *-cvs cases,
    adcc/tests/testcases.py:199-301): sigma and converged singlets against the oracle."""
    import adcc_b200 as adcc
    ref = adcc.ReferenceState(data, core_orbitals=core, frozen_core=frozen_core, frozen_virtual=frozen_virtual)
    matrix = adcc.AdcMatrix(method, ref)
    oref = oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core, frozen_core=frozen_core,
                                 frozen_virtual=frozen_virtual)
    omatrix = oracle.OracleAdcMatrix(method, oref)
    assert matrix.shape == omatrix.shape
    v, ov = random_vector(matrix, omatrix, 11)
    sig, osig = matrix.matvec(v), omatrix.matvec(ov)
    for k in osig:
        assert rel_err(sig[k].to_ndarray(), osig[k]) < SIGMA_RTOL, f"sigma {k}"
    state = adcc.run_adc(matrix, n_singlets=n, conv_tol=1e-8, output=None)
    ostate = oracle.run_adc(omatrix, method, n_singlets=n, conv_tol=1e-8)
    assert state.converged and ostate.converged and state.n_iter == ostate.n_iter
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    # properties with inactive orbitals: the frozen occupied orbitals stay part of the ground-state
    # density (the identity on o3, adcc/GroundState.py:137-146) - state dipole moments need them
    from oracle.properties import oscillator_strengths, state_dipole_moments
    of, _ = oscillator_strengths(omatrix, ostate)
    np.testing.assert_allclose(state.oscillator_strength, of, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(state.state_dipole_moment, state_dipole_moments(omatrix, ostate), rtol=0, atol=1e-6)


def check_extra_term(data):
    """AdcExtraTerm / matrix + term / term + matrix / matrix += term, following the reference's
    own test (adcc/tests/AdcMatrix_test.py:403-470): a shift added through the plug-in interface
    equals AdcMatrixShifted."""
    import pytest
    import adcc_b200 as adcc
    from adcc_b200 import AdcBlock, AdcExtraTerm, AdcMatrixShifted
    ground_state = adcc.LazyMp(adcc.ReferenceState(data))
    matrix_adc1 = adcc.AdcMatrix("adc1", ground_state)
    with pytest.raises(TypeError):
        matrix_adc1 += 42
    matrix = adcc.AdcMatrix("adc2", ground_state)
    with pytest.raises(TypeError):
        adcc.AdcMatrix("adc2", ground_state, diagonal_precomputed=42)
    with pytest.raises(TypeError):
        AdcExtraTerm(matrix, "fail")
    with pytest.raises(TypeError):
        AdcExtraTerm(matrix, {"fail": "not_callable"})
    shift = -0.3
    shifted = AdcMatrixShifted(matrix, shift)
    ones = matrix.diagonal().ones_like()

    def shift_ph(hf, mp, intermediates):
        return AdcBlock(lambda invec: adcc.AmplitudeVector(ph=shift * invec.ph),
                        adcc.AmplitudeVector(ph=shift * ones.ph))

    def shift_pphh(hf, mp, intermediates):
        return AdcBlock(lambda invec: adcc.AmplitudeVector(pphh=shift * invec.pphh),
                        adcc.AmplitudeVector(pphh=shift * ones.pphh))
    extra = AdcExtraTerm(matrix, {"ph_ph": shift_ph, "pphh_pphh": shift_pphh})
    with pytest.raises(ValueError):          # ADC(1) has no pphh_pphh block
        matrix_adc1 += extra
    omatrix = oracle.OracleAdcMatrix("adc2", oracle.OracleRefState(oracle.OracleHf(data)))
    vec, ovec = random_vector(matrix, omatrix, 21)
    ref = shifted @ vec
    osig = omatrix.matvec(ovec)
    for manual in (matrix + extra, extra + matrix):
        assert len(manual.extra_terms) == 1 and not matrix.extra_terms
        for b in ("ph", "pphh"):
            np.testing.assert_allclose(manual.diagonal()[b].to_ndarray(), shifted.diagonal()[b].to_ndarray(),
                                       atol=1e-12)
            np.testing.assert_allclose((manual @ vec)[b].to_ndarray(), ref[b].to_ndarray(), atol=1e-12)
            assert rel_err((manual @ vec)[b].to_ndarray(), osig[b] + shift * ovec[b]) < SIGMA_RTOL
    # lists of vectors go through the per-vector path when extra terms are present
    many = (matrix + extra) @ [vec, vec]
    np.testing.assert_allclose(many[1].pphh.to_ndarray(), ref.pphh.to_ndarray(), atol=1e-12)
    matrix += extra
    np.testing.assert_allclose((matrix @ vec).ph.to_ndarray(), ref.ph.to_ndarray(), atol=1e-12)
    state = adcc.run_adc(matrix, n_singlets=2, conv_tol=1e-8, output=None)
    plain = oracle.run_adc(data, "adc2", n_singlets=2, conv_tol=1e-8)
    np.testing.assert_allclose(state.excitation_energy, plain.eigenvalues + shift, atol=ENERGY_ATOL)


def check_dense_export(data, method, core, n_states):
    """AdcMatrix.to_ndarray in the antisymmetrised dense basis: symmetric, and its lowest distinct
    eigenvalues are the converged states -- the reference's own dense-export test
    (adcc/tests/AdcMatrix_dense_export_test.py:33-68)."""
    import adcc_b200 as adcc
    matrix = adcc.AdcMatrix(method, adcc.ReferenceState(data, core_orbitals=core))
    conv_tol = 1e-8
    state = adcc.run_adc(matrix, conv_tol=conv_tol, n_states=n_states, max_subspace=7 * n_states, output=None)
    dense = matrix.to_ndarray()
    basis = matrix.dense_basis()
    assert dense.shape == (len(basis), len(basis))
    np.testing.assert_allclose(dense, dense.T, rtol=1e-10, atol=1e-12)
    spectrum = np.linalg.eigvalsh(dense)
    rounded = np.unique(np.round(spectrum, 10))[:n_states]
    np.testing.assert_allclose(state.excitation_energy, rounded, atol=10 * conv_tol)
    return spectrum


def check_transfer_cvs_to_full(data, kind="singlet"):
    """transfer_cvs_to_full (adcc/projection.py:238-359): scalar products are preserved
    (adcc/tests/projection_test.py:223-251), and CVS eigenvectors serve as guesses for the
    projected full matrix (examples/2_core_excitations/water_core_by_projection.py:30-46)."""
    import pytest
    import adcc_b200 as adcc
    from adcc_b200.projection import transfer_cvs_to_full
    kw = {"n_singlets": 2} if kind == "singlet" else {"n_triplets": 2}
    state_cvs = adcc.run_adc(data, method="cvs-adc2x", core_orbitals=1, conv_tol=1e-8, output=None, **kw)
    matrix = adcc.AdcMatrix("adc2x", adcc.ReferenceState(data))
    gram = np.array([[v @ w for v in state_cvs.excitation_vector] for w in state_cvs.excitation_vector])
    full = transfer_cvs_to_full(state_cvs, matrix)
    gram_full = np.array([[v @ w for v in full] for w in full])
    np.testing.assert_allclose(gram, gram_full, atol=1e-14)
    for v in full:
        assert v.ph.subspaces == ["o1", "v1"] and v.pphh.subspaces == ["o1", "o1", "v1", "v1"]
        d = v.pphh.to_ndarray()
        assert np.max(np.abs(d + d.transpose(1, 0, 2, 3))) < 1e-15
    # random vectors of the CVS space, explicit kind
    rnd = [v.copy().set_random() for v in state_cvs.excitation_vector]
    gram = np.array([[v @ w for v in rnd] for w in rnd])
    full_rnd = transfer_cvs_to_full(state_cvs.matrix, matrix, rnd, kind)
    np.testing.assert_allclose(gram, np.array([[v @ w for v in full_rnd] for w in full_rnd]),
                               rtol=1e-12, atol=1e-12)      # random data: summation order differs
    with pytest.raises(ValueError):
        transfer_cvs_to_full(state_cvs.matrix, matrix)            # neither states nor vectors
    # the transferred eigenvectors as guesses of the projected matrix
    pm = adcc.AdcMatrixProjected(matrix, ["cv", "ccvv", "ocvv"], core_orbitals=1)
    by_guess = adcc.run_adc(pm, conv_tol=1e-8, guesses=transfer_cvs_to_full(state_cvs, pm), output=None, **kw)
    plain = adcc.run_adc(pm, conv_tol=1e-8, n_guesses=2, n_guesses_doubles=0, output=None, **kw)
    assert by_guess.converged and plain.converged
    np.testing.assert_allclose(by_guess.excitation_energy, plain.excitation_energy, atol=1e-7)
    assert by_guess.n_iter <= plain.n_iter
    assert np.max(np.abs(by_guess.excitation_energy - state_cvs.excitation_energy)) < 0.05


def check_davidson_edge_cases(data):
    """Solver control flow beyond the happy path (adcc/solver/davidson.py:225-245, 330-350,
    adcc/tests/solver/davidson_test.py): subspace collapse on a small max_subspace, the
    max_iter warning with converged=False, callbacks, and n_block > n_ep."""
    import warnings
    import pytest
    import adcc_b200 as adcc
    from adcc_b200.solver.davidson import jacobi_davidson
    from adcc_b200.guess import guesses_singlet
    from adcc_b200.solver.explicit_symmetrisation import IndexSpinSymmetrisation
    matrix, omatrix = build_pair(data, "adc2", None)
    # (1) restarts: max_subspace barely above the block size
    events, oevents = [], []
    state = adcc.run_adc(matrix, n_singlets=3, conv_tol=1e-8, max_subspace=7, output=None)
    ostate = oracle.run_adc(omatrix, "adc2", n_singlets=3, conv_tol=1e-8, max_subspace=7,
                            callback=lambda st, ident: oevents.append(ident))
    direct = jacobi_davidson(matrix, guesses_singlet(matrix, 6, block="ph"), n_ep=3, conv_tol=1e-8,
                             max_subspace=7, callback=lambda st, ident: events.append(ident),
                             explicit_symmetrisation=IndexSpinSymmetrisation(matrix, "singlet"))
    assert events[0] == "start" and events[-1] == "is_converged" and events.count("restart") >= 3
    assert events.count("next_iter") == oevents.count("next_iter") == state.n_iter
    assert direct.n_iter == state.n_iter
    with pytest.raises(TypeError):            # like the reference: run_adc owns the solver callback
        adcc.run_adc(matrix, n_singlets=3, output=None, callback=lambda st, ident: None)
    assert state.converged and state.n_iter == ostate.n_iter and state.n_applies == ostate.n_applies
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    # (2) too few iterations: warning, converged False, best estimates returned
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        short = adcc.run_adc(matrix, n_singlets=3, conv_tol=1e-10, max_iter=2, output=None)
    assert not short.converged and short.n_iter == 2
    assert any("Maximum number of iterations" in str(w.message) for w in caught)
    assert np.all(np.abs(short.excitation_energy - state.excitation_energy) < 1e-2)
    # (3) more guesses than states and a block size of its own through the solver interface
    guesses = guesses_singlet(matrix, 6, block="ph")
    res = jacobi_davidson(matrix, guesses, n_ep=2, n_block=4, conv_tol=1e-8,
                          explicit_symmetrisation=IndexSpinSymmetrisation(matrix, "singlet"))
    assert res.converged and len(res.eigenvalues) == 2
    np.testing.assert_allclose(res.eigenvalues, state.excitation_energy[:2], atol=1e-7)
    # (4) invalid solver arguments
    from adcc_b200.solver.davidson import eigsh
    with pytest.raises(ValueError):
        eigsh(matrix, guesses, n_ep=2, conv_tol=1e-8, preconditioning_method="nonsense")
    with pytest.raises(ValueError):
        eigsh(matrix, guesses[:1], n_ep=2, conv_tol=1e-8)            # fewer guesses than states
