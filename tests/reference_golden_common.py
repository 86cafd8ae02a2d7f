"""Comparisons against tests/golden/reference_adcc_h2o_sto3g.npz: results the REFERENCE's own,
unmodified Python layer (adcc/adc_pp/matrix.py, LazyMp.py, solver/davidson.py ...) produced when it
was run on this repository's ``libadcc`` surface (see tests/golden/make_reference_adcc_goldens.py)."""
import os

import numpy as np

from oracle.hfdata import load_h2o_sto3g

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_adcc_h2o_sto3g.npz")
METHODS = ["adc0", "adc1", "adc2", "adc2x", "adc3", "cvs-adc0", "cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3"]
SIGMA_RTOL = 1e-11        # north star; the reference's own test uses rtol 1e-10 (AdcMatrix_test.py:221)
ENERGY_ATOL = 1e-8
OSC_RTOL, OSC_ATOL = 1e-6, 1e-10    # "matching oscillator strengths": 1e-6 relative
# state densities are quadratic in eigenvectors converged to a residual norm of 1e-9
DIPOLE_ATOL, DENSITY_ATOL = 1e-7, 1e-7
_SPACES = {"o": "o1", "v": "v1", "c": "o2"}


def golden():
    return np.load(GOLDEN)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def blocks_of(g, key):
    return sorted(k.split("/")[-1] for k in g.files if k.startswith(f"{key}/vector/"))


def check_oracle(method):
    import oracle
    g = golden()
    key = method.replace("-", "_")
    core = 1 if method.startswith("cvs") else None
    data = load_h2o_sto3g()
    om = oracle.OracleAdcMatrix(method, oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core))
    v = {b: g[f"{key}/vector/{b}"] for b in blocks_of(g, key)}
    sig = om.matvec(v)
    for b in v:
        assert rel(sig[b], g[f"{key}/matvec/{b}"]) < 1e-12, (method, b)
        assert rel(om.diagonal()[b], g[f"{key}/diagonal/{b}"]) < 1e-12, (method, "diagonal", b)
    for blk in om.blocks:
        if f"{key}/block/{blk}" not in g.files:
            continue
        outb, inb = blk.split("_")
        res = om.blocks[blk]({inb: v[inb]})
        want = g[f"{key}/block/{blk}"]
        got = res[outb] if res is not None else np.zeros_like(want)
        assert rel(got, want) < 1e-12 or np.max(np.abs(want)) == 0.0, (method, blk)
    if f"{key}/mp2" in g.files:
        assert abs(om.mp.energy_correction(2) - float(g[f"{key}/mp2"])) < 1e-12
    if f"{key}/mp3" in g.files:
        assert abs(om.mp.energy_correction(3) - float(g[f"{key}/mp3"])) < 1e-12
    for kind in ("singlet", "triplet"):
        want = g[f"{key}/{kind}/energies"]
        st = oracle.run_adc(data, method, conv_tol=1e-9, core_orbitals=core,
                            **{"n_singlets" if kind == "singlet" else "n_triplets": len(want)})
        np.testing.assert_allclose(st.eigenvalues, want, atol=ENERGY_ATOL, rtol=0)
        if kind == "singlet" and f"{key}/singlet/oscillator_strength" in g.files:
            # the reference's own property code (adc_pp/transition_dm.py, OperatorIntegrals.py) produced these
            from oracle.properties import oscillator_strengths
            f, mu = oscillator_strengths(om, st)
            np.testing.assert_allclose(f, g[f"{key}/singlet/oscillator_strength"], rtol=OSC_RTOL, atol=OSC_ATOL)
            np.testing.assert_allclose(np.abs(mu), np.abs(g[f"{key}/singlet/transition_dipole_moment"]),
                                       rtol=0, atol=1e-6)
        if kind == "singlet":
            # state densities of the reference's own property code (adc_pp/state_diffdm.py,
            # ElectronicStates.state_dipole_moment, GroundState.dipole_moment); quadratic in the
            # eigenvector, so its sign does not matter
            from oracle.properties import ground_state_dipole_moment, state_diffdm, state_dipole_moments
            level = min(int((method[4:] if core else method)[3]), 2)
            np.testing.assert_allclose(ground_state_dipole_moment(om, level), g[f"{key}/ground_state_dipole_moment"],
                                       rtol=0, atol=1e-10)
            np.testing.assert_allclose(state_dipole_moments(om, st), g[f"{key}/singlet/state_dipole_moment"],
                                       rtol=0, atol=DIPOLE_ATOL)
            for n, vec in enumerate(st.eigenvectors):
                dm = state_diffdm(om, vec)
                for blk, rho in dm.items():
                    gk = f"{key}/singlet/state_diffdm/{n}/{_SPACES[blk[0]]}{_SPACES[blk[1]]}"
                    if gk in g.files:
                        assert np.max(np.abs(rho - g[gk])) < DENSITY_ATOL, (method, n, blk)
                    else:                                   # the reference dropped it as a zero block
                        assert np.max(np.abs(rho)) < DENSITY_ATOL, (method, n, blk)
            if f"{key}/singlet/s2s/excitation_energy" in g.files:
                from oracle.properties import state2state
                _compare_s2s(g, key, *state2state(om, st, initial=0), to_array=lambda a: a)


def _compare_s2s(g, key, ene, mus, f, dms, to_array):
    """State-to-state quantities against the reference's State2States output; the density and the
    dipole carry the product of two eigenvector signs, so they are compared up to one sign each."""
    np.testing.assert_allclose(ene, g[f"{key}/singlet/s2s/excitation_energy"], rtol=0, atol=ENERGY_ATOL)
    np.testing.assert_allclose(f, g[f"{key}/singlet/s2s/oscillator_strength"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(np.abs(mus), np.abs(g[f"{key}/singlet/s2s/transition_dipole_moment"]),
                               rtol=0, atol=DIPOLE_ATOL)
    for n, dm in enumerate(dms):
        got, want = {}, {}
        for blk, rho in dm.items():
            got[blk] = to_array(rho)
            gk = f"{key}/singlet/s2s/transition_dm/{n}/{_SPACES[blk[0]]}{_SPACES[blk[1]]}"
            want[blk] = g[gk] if gk in g.files else np.zeros_like(got[blk])      # dropped as a zero block
        big = max(want, key=lambda blk: np.max(np.abs(want[blk])))              # the sign from the largest element
        k = np.unravel_index(np.argmax(np.abs(want[big])), want[big].shape)
        sign = 1.0 if got[big][k] * want[big][k] >= 0 else -1.0
        for blk in got:
            assert np.max(np.abs(sign * got[blk] - want[blk])) < DENSITY_ATOL, (key, n, blk)


def check_product(method, doubles_storage="dense"):
    """The product path (whatever backend is active: CUDA on the GPU box, the CPU stand-in in the
    build container) against the reference-produced goldens."""
    import adcc_b200 as adcc
    g = golden()
    key = method.replace("-", "_")
    core = 1 if method.startswith("cvs") else None
    data = load_h2o_sto3g()
    ref = adcc.ReferenceState(data, core_orbitals=core)
    m = adcc.AdcMatrix(method, ref, doubles_storage=doubles_storage)
    bl = blocks_of(g, key)

    def make(b):
        t = adcc.Tensor.from_ndarray(m.mospaces, m.axis_spaces[b], g[f"{key}/vector/{b}"])
        if doubles_storage == "packed" and b == "pphh":
            from adcc_b200.packed import PackedDoubles
            from adcc_b200.packed_cvs import CvsPackedDoubles
            t = (CvsPackedDoubles if core else PackedDoubles).from_dense(t)
        return t
    v = adcc.AmplitudeVector(**{b: make(b) for b in bl})
    sig = m.matvec(v)
    for b in bl:
        assert rel(sig[b].to_ndarray(), g[f"{key}/matvec/{b}"]) < SIGMA_RTOL, (method, b)
        d, want = m.diagonal()[b].to_ndarray(), g[f"{key}/diagonal/{b}"]
        if doubles_storage == "packed" and b == "pphh":        # i == j / a == b entries are not stored
            no, nv = want.shape[0], want.shape[2]
            mask = np.ones_like(want, dtype=bool)
            if not core:                                       # CVS doubles u[i,J,a,b]: i and J are different spaces
                mask[np.arange(no), np.arange(no)] = False
            mask[:, :, np.arange(nv), np.arange(nv)] = False
            d, want = d[mask], want[mask]
        assert rel(d, want) < SIGMA_RTOL, (method, "diagonal", b)
    for blk in m.blocks:
        if f"{key}/block/{blk}" not in g.files:
            continue
        outb, inb = blk.split("_")
        want = g[f"{key}/block/{blk}"]
        if np.max(np.abs(want)) == 0.0:
            continue
        res = m.block_apply(blk, v[inb])
        assert rel(res.to_ndarray(), want) < SIGMA_RTOL, (method, blk)
    if doubles_storage == "dense":
        if f"{key}/mp2" in g.files:
            assert abs(m.ground_state.energy_correction(2) - float(g[f"{key}/mp2"])) < 1e-12
        if f"{key}/mp3" in g.files:
            assert abs(m.ground_state.energy_correction(3) - float(g[f"{key}/mp3"])) < 1e-12
    for kind in ("singlet", "triplet"):
        want = g[f"{key}/{kind}/energies"]
        st = adcc.run_adc(m, conv_tol=1e-9, output=None,
                          **{"n_singlets" if kind == "singlet" else "n_triplets": len(want)})
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, want, atol=ENERGY_ATOL, rtol=0)
        assert st.n_iter == int(g[f"{key}/{kind}/n_iter"]), (method, kind, st.n_iter)
        if kind == "singlet" and f"{key}/singlet/oscillator_strength" in g.files:
            np.testing.assert_allclose(st.oscillator_strength, g[f"{key}/singlet/oscillator_strength"],
                                       rtol=OSC_RTOL, atol=OSC_ATOL)
            np.testing.assert_allclose(np.abs(st.transition_dipole_moment),
                                       np.abs(g[f"{key}/singlet/transition_dipole_moment"]), rtol=0, atol=1e-6)
        if kind == "singlet" and doubles_storage == "dense":
            level = min(int((method[4:] if core else method)[3]), 2)
            np.testing.assert_allclose(m.ground_state.dipole_moment(level), g[f"{key}/ground_state_dipole_moment"],
                                       rtol=0, atol=1e-10)
            np.testing.assert_allclose(st.state_dipole_moment, g[f"{key}/singlet/state_dipole_moment"],
                                       rtol=0, atol=DIPOLE_ATOL)
            for n, dm in enumerate(st.state_diffdm):
                for blk, rho in dm.items():
                    gk = f"{key}/singlet/state_diffdm/{n}/{_SPACES[blk[0]]}{_SPACES[blk[1]]}"
                    if gk in g.files:
                        assert np.max(np.abs(rho.to_ndarray() - g[gk])) < DENSITY_ATOL, (method, n, blk)
                    else:
                        assert np.max(np.abs(rho.to_ndarray())) < DENSITY_ATOL, (method, n, blk)
            if f"{key}/singlet/s2s/excitation_energy" in g.files:
                s2s = adcc.State2States(st, initial=0)
                _compare_s2s(g, key, s2s.excitation_energy, s2s.transition_dipole_moment, s2s.oscillator_strength,
                             s2s.transition_dm, to_array=lambda t: t.to_ndarray())
        elif kind == "singlet":                              # packed doubles: same numbers through to_dense()
            np.testing.assert_allclose(st.state_dipole_moment, g[f"{key}/singlet/state_dipole_moment"],
                                       rtol=0, atol=DIPOLE_ATOL)


LANCZOS_METHODS = ["adc2", "adc2x", "cvs-adc2"]


def check_lanczos(method, doubles_storage="dense", iter_slack=0):
    """run_adc(..., eigensolver="lanczos") of the product against what the reference's own Lanczos
    solver (adcc/solver/lanczos.py, LanczosIterator.py, orthogonaliser.py) produced on the libadcc
    surface: energies to 1e-8 and identical iteration counts."""
    import warnings
    import adcc_b200 as adcc
    g = golden()
    key = method.replace("-", "_")
    core = 1 if method.startswith("cvs") else None
    ref = adcc.ReferenceState(load_h2o_sto3g(), core_orbitals=core)
    m = adcc.AdcMatrix(method, ref, doubles_storage=doubles_storage)
    for kind in ("singlet", "triplet"):
        want = g[f"{key}/lanczos/{kind}/energies"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = adcc.run_adc(m, conv_tol=1e-8, output=None, eigensolver="lanczos",
                              **{"n_singlets" if kind == "singlet" else "n_triplets": len(want)})
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, want, atol=ENERGY_ATOL, rtol=0)
        # 30 - 70 steps whose last one is decided by |r| |b^T y| against 1e-8 |theta|: device rounding may
        # move it by a step (iter_slack), the host path reproduces the count
        assert abs(st.n_iter - int(g[f"{key}/lanczos/{kind}/n_iter"])) <= iter_slack, (method, kind, st.n_iter)
        # the same states as the Jacobi-Davidson solver finds
        np.testing.assert_allclose(st.excitation_energy, g[f"{key}/{kind}/energies"], atol=1e-7, rtol=0)
        assert np.all(st.solver_state.residual_norms < 1e-5)


OPEN_SHELL_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_adcc_open_shell.npz")
OPEN_SHELL_METHODS = ["adc1", "adc2", "adc2x", "adc3"]


def check_open_shell(method, which="product", iter_slack=0):
    """The unrestricted open-shell input (tests/adc_parity_common.py::open_shell_dict) against what
    the reference's own Python layer computed for it: sigma, diagonal, MP2 energy, kind "any" and
    "spin_flip" states with the reference's iteration counts."""
    import warnings
    from adc_parity_common import open_shell_dict
    g = np.load(OPEN_SHELL_GOLDEN)
    data = open_shell_dict()
    blocks = sorted(k.split("/")[-1] for k in g.files if k.startswith(f"{method}/vector/"))
    if which == "oracle":
        import oracle
        om = oracle.OracleAdcMatrix(method, oracle.OracleRefState(oracle.OracleHf(data)))
        sig = om.matvec({b: g[f"{method}/vector/{b}"] for b in blocks})
        for b in blocks:
            assert rel(sig[b], g[f"{method}/matvec/{b}"]) < 1e-12, (method, b)
            assert rel(om.diagonal()[b], g[f"{method}/diagonal/{b}"]) < 1e-12, (method, "diagonal", b)
        assert abs(om.mp.energy_correction(2) - float(g[f"{method}/mp2"])) < 1e-12
        for kind, n in (("any", 3), ("spin_flip", 2)):
            st = oracle.run_adc(om, method, n_states=n, kind=kind, conv_tol=1e-8)
            np.testing.assert_allclose(st.eigenvalues, g[f"{method}/{kind}/energies"], atol=ENERGY_ATOL, rtol=0)
        return
    import adcc_b200 as adcc
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="dense")
    v = adcc.AmplitudeVector(**{b: adcc.Tensor.from_ndarray(m.mospaces, m.axis_spaces[b], g[f"{method}/vector/{b}"])
                                for b in blocks})
    sig = m.matvec(v)
    for b in blocks:
        assert rel(sig[b].to_ndarray(), g[f"{method}/matvec/{b}"]) < SIGMA_RTOL, (method, b)
        assert rel(m.diagonal()[b].to_ndarray(), g[f"{method}/diagonal/{b}"]) < SIGMA_RTOL, (method, "diagonal", b)
    assert abs(m.ground_state.energy_correction(2) - float(g[f"{method}/mp2"])) < 1e-12
    for kind, kw, n in (("any", "n_states", 3), ("spin_flip", "n_spin_flip", 2)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = adcc.run_adc(m, conv_tol=1e-8, output=None, **{kw: n})
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, g[f"{method}/{kind}/energies"], atol=ENERGY_ATOL, rtol=0)
        assert abs(st.n_iter - int(g[f"{method}/{kind}/n_iter"])) <= iter_slack, (method, kind, st.n_iter)


WORKFLOW_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_adcc_workflows.npz")


def check_workflows():
    """Callers around the matvec on H2O / STO-3G against the reference's own runs: the projected
    ADC(2)-x matrix of examples/2_core_excitations/water_core_by_projection.py (AdcMatrixProjected,
    CVS guesses transferred by transfer_cvs_to_full) and the frozen-core / frozen-virtual cases of
    adcc/tests/testcases.py (fc, fv, fc-fv, fv-cvs)."""
    import warnings
    import adcc_b200 as adcc
    from adcc_b200.projection import transfer_cvs_to_full
    g = np.load(WORKFLOW_GOLDEN)
    data = load_h2o_sto3g()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        matrix = adcc.AdcMatrix("adc2x", adcc.ReferenceState(data))
        pmatrix = adcc.AdcMatrixProjected(matrix, ["cv", "ccvv", "ocvv"], core_orbitals=1)
        sp = adcc.run_adc(pmatrix, n_singlets=2, conv_tol=1e-8, n_guesses=2, n_guesses_doubles=0, output=None)
        sc = adcc.cvs_adc2x(data, n_singlets=2, conv_tol=1e-8, core_orbitals=1, output=None)
        sp2 = adcc.run_adc(pmatrix, n_singlets=2, conv_tol=1e-8, guesses=transfer_cvs_to_full(sc, pmatrix),
                           output=None)
    for st, key in ((sp, "projected"), (sp2, "projected_from_cvs")):
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, g[f"{key}/energies"], atol=ENERGY_ATOL, rtol=0)
        assert st.n_iter == int(g[f"{key}/n_iter"]), (key, st.n_iter)
    for blk in ("ph", "pphh"):
        assert rel(pmatrix.diagonal()[blk].to_ndarray(), g[f"projected/diagonal/{blk}"]) < SIGMA_RTOL
    for name, method, kw in (("fc", "adc2", dict(frozen_core=1)), ("fv", "adc2", dict(frozen_virtual=1)),
                             ("fc_fv", "adc2", dict(frozen_core=1, frozen_virtual=1)),
                             ("fv_cvs", "cvs-adc2", dict(core_orbitals=1, frozen_virtual=1)),
                             ("fc_adc3", "adc3", dict(frozen_core=1))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = adcc.run_adc(data, method=method, n_singlets=2, conv_tol=1e-8, output=None, **kw)
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, g[f"{name}/energies"], atol=ENERGY_ATOL, rtol=0)
        assert st.n_iter == int(g[f"{name}/n_iter"]), (name, st.n_iter)
        if f"{name}/oscillator_strength" in g.files:
            np.testing.assert_allclose(st.oscillator_strength, g[f"{name}/oscillator_strength"],
                                       rtol=OSC_RTOL, atol=OSC_ATOL)


BASELINE_SHAPES_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                      "reference_adcc_baseline_shapes.npz")
BASELINE_SHAPE_CASES = {     # as in tests/golden/make_reference_adcc_goldens.py
    "c1": ("h2o_ccpvdz", "adc2", dict(n_singlets=3, conv_tol=1e-6)),
    "c1_adc3": ("h2o_ccpvdz", "adc3", dict(n_singlets=3, conv_tol=1e-6)),
    "c2": ("h2o_ccpvtz", "adc2", dict(n_singlets=10)),
    "c4": ("h2o_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=7, conv_tol=1e-8)),
    "c4_ne": ("ne_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=4, conv_tol=1e-8)),
    "c3": ("methyloxirane_ccpvdz", "adc3", dict(n_singlets=8)),
}


def check_baseline_shape(key, run=True, doubles_storage="auto"):
    """The restricted synthetic SCF input of one exact BASELINE shape (configs[0..3]) against the
    reference's OWN run_adc on it (through the libadcc surface): sigma of a seeded trial vector
    (singles complete, 4096 sampled doubles) to 1e-11, then - ``run`` - energies within the
    convergence tolerance, oscillator strengths and the iteration count."""
    import warnings
    import pytest
    import adcc_b200 as adcc
    from adcc_b200.synthetic import named_case
    g = np.load(BASELINE_SHAPES_GOLDEN)
    if f"{key}/energies" not in g.files:
        pytest.skip(f"no reference-produced golden for {key} (hours of CPU; see the generator)")
    name, method, kw = BASELINE_SHAPE_CASES[key]
    data, core = named_case(name)
    ref = adcc.ReferenceState(data, core_orbitals=core or None)
    m = adcc.AdcMatrix(method, ref, doubles_storage=doubles_storage)
    cvs = method.startswith("cvs")
    rng = np.random.default_rng(int(g[f"{key}/vector_seed"]))
    blocks = {}
    for blk in ("ph", "pphh"):                        # the generator's order of draws
        arr = rng.standard_normal(tuple(m.mospaces.n_orbs(sp) for sp in m.axis_spaces[blk]))
        if blk == "pphh":
            arr = arr - arr.transpose(0, 1, 3, 2)
            if not cvs:
                arr = arr - arr.transpose(1, 0, 2, 3)
        t = adcc.Tensor.from_ndarray(m.mospaces, m.axis_spaces[blk], arr)
        if blk == "pphh" and m.doubles_storage == "packed":
            from adcc_b200.packed import PackedDoubles
            from adcc_b200.packed_cvs import CvsPackedDoubles
            t = (CvsPackedDoubles if cvs else PackedDoubles).from_dense(t)
        blocks[blk] = t
    sig = m.matvec(adcc.AmplitudeVector(**blocks))
    assert rel(sig.ph.to_ndarray(), g[f"{key}/matvec/ph"]) < SIGMA_RTOL, key
    s2 = sig.pphh.to_ndarray().reshape(-1)[g[f"{key}/matvec/pphh_index"]]
    assert np.max(np.abs(s2 - g[f"{key}/matvec/pphh_value"])) < SIGMA_RTOL * float(g[f"{key}/matvec/pphh_max"]), key
    if not run:
        return
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = adcc.run_adc(m, output=None, **kw)
    tol = kw.get("conv_tol", 1e-6)
    assert st.converged
    np.testing.assert_allclose(st.excitation_energy, g[f"{key}/energies"], atol=max(10 * tol ** 2, 1e-8) + tol * 1e-2, rtol=0)
    np.testing.assert_allclose(st.oscillator_strength, g[f"{key}/oscillator_strength"], rtol=1e-3, atol=1e-4 * tol / 1e-6)
    assert st.n_iter == int(g[f"{key}/n_iter"]), (key, st.n_iter)


SYNTHETIC_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                "reference_adcc_synthetic_workload.npz")


def check_synthetic_workload(no, nv, seed, method):
    """BASELINE configs[4]'s input definition at a small size.  Reference side (golden): the counter-based
    <pq||rs> written out as a plain SCF dict and run through the reference's own ReferenceState /
    AdcMatrix / run_adc.  Product side: the device generator + pair-packed store + fused engine that
    the headline benchmark uses (synthetic_reference_state, doubles_storage="packed").  Sigma and
    diagonal to 1e-11, three states to 1e-8 with the reference's iteration count."""
    import warnings
    import adcc_b200 as adcc
    from adcc_b200.packed import PackedDoubles, synthetic_reference_state
    g = np.load(SYNTHETIC_GOLDEN)
    key = f"{no}_{nv}/{method}"
    ref = synthetic_reference_state(no, nv, seed=seed, scale=0.02)
    m = adcc.AdcMatrix(method, ref, doubles_storage="packed")
    ph = adcc.Tensor.from_ndarray(m.mospaces, m.axis_spaces["ph"], g[f"{key}/vector/ph"])
    pphh = PackedDoubles.from_dense(adcc.Tensor.from_ndarray(m.mospaces, m.axis_spaces["pphh"],
                                                             g[f"{key}/vector/pphh"]))
    sig = m.matvec(adcc.AmplitudeVector(ph=ph, pphh=pphh))
    for blk in ("ph", "pphh"):
        assert rel(sig[blk].to_ndarray(), g[f"{key}/matvec/{blk}"]) < SIGMA_RTOL, (key, blk)
        d, want = m.diagonal()[blk].to_ndarray(), g[f"{key}/diagonal/{blk}"]
        if blk == "pphh":                              # i == j / a == b entries are not stored
            mask = np.ones_like(want, dtype=bool)
            mask[np.arange(no), np.arange(no)] = False
            mask[:, :, np.arange(nv), np.arange(nv)] = False
            d, want = d[mask], want[mask]
        assert rel(d, want) < SIGMA_RTOL, (key, "diagonal", blk)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = adcc.run_adc(m, n_states=3, conv_tol=1e-8, output=None)
    assert st.converged
    np.testing.assert_allclose(st.excitation_energy, g[f"{key}/energies"], atol=ENERGY_ATOL, rtol=0)
    assert st.n_iter == int(g[f"{key}/n_iter"]), (key, st.n_iter)


def check_synthetic_workload_oracle(no, nv, seed, method):
    """The dense oracle on the synthetic input definition (oracle.synthetic.OracleSyntheticRef) and -
    for ADC(2)-x - the full-size restatement that serves as the CPU arm of bench.py
    (oracle/headline.py) against the reference's own run on the same integrals.  This closes the chain
    reference code -> dense oracle -> headline oracle -> whole GPU sigma at nocc=40 / nvirt=400."""
    import oracle
    from oracle.synthetic import OracleSyntheticRef
    g = np.load(SYNTHETIC_GOLDEN)
    key = f"{no}_{nv}/{method}"
    om = oracle.OracleAdcMatrix(method, OracleSyntheticRef(no, nv, seed=seed, scale=0.02))
    v = {b: g[f"{key}/vector/{b}"] for b in ("ph", "pphh")}
    sig = om.matvec(v)
    for b in v:
        assert rel(sig[b], g[f"{key}/matvec/{b}"]) < 1e-12, (key, b)
        assert rel(om.diagonal()[b], g[f"{key}/diagonal/{b}"]) < 1e-12, (key, "diagonal", b)
    st = oracle.run_adc(om, method, n_states=3, conv_tol=1e-8)
    np.testing.assert_allclose(st.eigenvalues, g[f"{key}/energies"], atol=ENERGY_ATOL, rtol=0)
    assert st.n_iter == int(g[f"{key}/n_iter"])
    if method == "adc2x":
        from oracle.headline import OracleHeadline
        oh = OracleHeadline(no, nv, seed=seed, scale=0.02)
        po = [(i, j) for j in range(no) for i in range(j)]
        pv = [(a, b) for b in range(nv) for a in range(b)]
        u2 = v["pphh"]
        packed = np.array([[u2[i, j, a, b] for (a, b) in pv] for (i, j) in po])
        s1, s2 = oh.matvec(v["ph"], packed)
        want2 = np.array([[g[f"{key}/matvec/pphh"][i, j, a, b] for (a, b) in pv] for (i, j) in po])
        assert rel(s1, g[f"{key}/matvec/ph"]) < 1e-12 and rel(s2, want2) < 1e-12, key
