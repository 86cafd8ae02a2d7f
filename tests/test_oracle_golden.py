"""Pin the CPU oracle against the reference's in-source known answers.

Known answers: adcc/tests/smoke_test.py:32-58 (H2O/STO-3G, atol 1e-6); fixture:
tests/golden/h2o_sto3g.npz generated from examples/water/data.py by
tests/golden/make_h2o_sto3g.py.
"""
import numpy as np
import pytest

import oracle
from oracle.hfdata import load_h2o_sto3g

SINGLETS = {
    "adc0": [0.99048446, 1.04450741, 1.15356339],
    "adc1": [0.48343609, 0.57420044, 0.60213699],
    "adc2": [0.47051314, 0.57255495, 0.59367339],
    "adc2x": [0.44511182, 0.55139508, 0.57146893],
    "adc3": [0.45197068, 0.55572706, 0.57861241],
    "cvs-adc0": [20.83623681, 20.99931574],
    "cvs-adc1": [20.10577515, 20.16512502],
    "cvs-adc2": [20.0045422, 20.08771799],
    "cvs-adc2x": [19.90528006, 19.9990468],
    "cvs-adc3": [19.93138526, 20.02055267],
}
TRIPLETS = {
    "adc0": [0.99048446, 1.04450741, 1.15356339],
    "adc1": [0.40544164, 0.48351268, 0.52489129],
    "adc2": [0.40288477, 0.4913253, 0.52854722],
    "adc2x": [0.38557787, 0.48177562, 0.51336301],
    "adc3": [0.3924605, 0.4877098, 0.51755258],
    "cvs-adc0": [20.83623681, 20.99931574],
    "cvs-adc1": [20.04302084, 20.12099631],
    "cvs-adc2": [19.95732865, 20.05239094],
    "cvs-adc2x": [19.86572075, 19.97241019],
    "cvs-adc3": [19.89172578, 19.99386631],
}


@pytest.fixture(scope="module")
def h2o():
    return load_h2o_sto3g()


@pytest.mark.parametrize("method", sorted(SINGLETS))
@pytest.mark.parametrize("kind", ["singlet", "triplet"])
def test_smoke_table(h2o, method, kind):
    ref = (SINGLETS if kind == "singlet" else TRIPLETS)[method]
    kw = {"n_singlets": len(ref)} if kind == "singlet" else {"n_triplets": len(ref)}
    core = 1 if method.startswith("cvs") else None
    state = oracle.run_adc(h2o, method, conv_tol=1e-8, core_orbitals=core, **kw)
    assert state.converged
    assert state.n_iter <= 40     # adcc/tests/functionality_test.py:135
    np.testing.assert_allclose(state.eigenvalues, ref, atol=1e-6)


def test_adc0_one_iteration(h2o):
    # adcc/tests/solver/davidson_test.py:92-94
    state = oracle.run_adc(h2o, "adc0", n_singlets=3, conv_tol=1e-8)
    assert state.n_iter == 1


def test_mp_energies(h2o):
    # survey-derived anchors
    ref = oracle.OracleRefState(oracle.OracleHf(h2o))
    mp = oracle.OracleMp(ref)
    assert abs(mp.energy_correction(2) - (-0.034258801922)) < 1e-10
    assert abs(mp.energy_correction(3) - (-0.009296984621)) < 1e-10


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3", "cvs-adc2x", "cvs-adc3"])
def test_matvec_hermitian_and_symmetric(h2o, method):
    core = 1 if method.startswith("cvs") else None
    ref = oracle.OracleRefState(oracle.OracleHf(h2o), core_orbitals=core)
    m = oracle.OracleAdcMatrix(method, ref)
    rng = np.random.default_rng(1)

    def rand():
        v = {k: rng.standard_normal(s) for k, s in m.shapes.items()}
        v["pphh"] = m.symmetrise_doubles(v["pphh"])
        return v
    u, w = rand(), rand()
    mu, mw = m.matvec(u), m.matvec(w)
    dot = lambda a, b: sum(float(np.vdot(a[k], b[k])) for k in a)  # noqa: E731
    assert abs(dot(u, mw) - dot(mu, w)) < 1e-11 * max(1.0, abs(dot(u, mw)))
    # sigma_pphh keeps the doubles symmetry
    np.testing.assert_allclose(m.symmetrise_doubles(mu["pphh"]), mu["pphh"], atol=1e-13)


def test_singlet_fixed_point_of_spin_enforcement(h2o):
    # converged singlets are fixed points of a27
    from oracle.davidson import enforce_spin_kind
    state = oracle.run_adc(h2o, "adc2", n_singlets=1, conv_tol=1e-9)
    u2 = state.eigenvectors[0]["pphh"]
    before = u2.copy()
    enforce_spin_kind(u2, [5, 5, 2, 2], "singlet")
    np.testing.assert_allclose(u2, before, atol=1e-9)


def test_oracle_symmetrisers_against_reference_goldens():
    """oracle.mp.sym / asym / sym_pairs vs libadcc_src/tests/TensorTestData.cc."""
    from oracle.mp import sym, asym, sym_pairs
    from tensor_golden_common import GOLD
    a = GOLD["a"]
    np.testing.assert_allclose(sym(a, 0, 1), GOLD["a_sym_01"], atol=1e-15)
    np.testing.assert_allclose(asym(a, 0, 1), GOLD["a_asym_01"], atol=1e-15)
    np.testing.assert_allclose(sym_pairs(a, (0, 1), (2, 3)), GOLD["a_sym_01_23"], atol=1e-15)
    np.testing.assert_allclose(0.5 * (a - a.transpose(1, 0, 3, 2)), GOLD["a_asym_01_23"], atol=1e-15)


def test_oracle_unique_pairs_mode_is_the_same_matrix():
    """The CPU timing baseline evaluates the oooo / vvvv terms over canonical index pairs (like
    libtensor's symmetry-unique blocks): identical sigma, and the slab sampler returns sane numbers."""
    import numpy as np
    import oracle
    from oracle.adc import dominant_terms_sample
    from oracle.synthetic import OracleSyntheticRef
    ref = OracleSyntheticRef(6, 10, seed=3, scale=0.02)
    dense = oracle.OracleAdcMatrix("adc2x", ref)
    pairs = oracle.OracleAdcMatrix("adc2x", ref, unique_pairs=True)
    rng = np.random.default_rng(1)
    v = {k: rng.standard_normal(s) for k, s in dense.shapes.items()}
    v["pphh"] = dense.symmetrise_doubles(v["pphh"])
    sd, sp = dense.matvec(v), pairs.matvec(v)
    for k in sd:
        assert np.max(np.abs(sd[k] - sp[k])) < 1e-13 * np.max(np.abs(sd[k]))
    flops, seconds = dominant_terms_sample(6, 10, 16, 16)
    assert flops == 2.0 * 15 * 16 * 45 + 2.0 * 60 * 16 * 60 and seconds > 0
