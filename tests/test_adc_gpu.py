"""Parity of the CUDA product path (through the C ABI) against the CPU oracle.

sigma vectors: max|delta| / max|sigma| <= 1e-11; excitation energies <= 1e-8 Ha (north star).
"""
import numpy as np
import pytest

from oracle.hfdata import load_h2o_sto3g
from adc_parity_common import (METHODS, SINGLETS, check_matvec_parity, check_run_parity)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h2o():
    return load_h2o_sto3g()


@pytest.fixture(autouse=True)
def _native_loaded():
    from adcc_b200 import _cabi
    assert _cabi.lib().adccb_version() >= 100      # the CUDA library is the thing under test
    before = _cabi.launch_count()
    yield
    assert _cabi.launch_count() > before           # kernels of libadccb.so really ran


@pytest.mark.parametrize("method", METHODS)
def test_matvec_vs_oracle_h2o(h2o, method):
    check_matvec_parity(h2o, method, 1 if method.startswith("cvs") else None)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("kind", ["singlet", "triplet"])
def test_run_adc_vs_oracle_and_smoke_table(h2o, method, kind):
    state, _ = check_run_parity(h2o, method, kind, 1 if method.startswith("cvs") else None)
    if kind == "singlet":
        # reference known answers, adcc/tests/smoke_test.py:32-58
        np.testing.assert_allclose(state.excitation_energy, SINGLETS[method], atol=1e-6)
    assert state.n_iter <= 40          # adcc/tests/functionality_test.py:135


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_matvec_synthetic_c1_shape(method):
    # BASELINE configs[0] shape: o=10, v=38 spin orbitals
    from adcc_b200.synthetic import named_case
    data, core = named_case("h2o_ccpvdz")
    check_matvec_parity(data, method, None)


@pytest.mark.parametrize("method", ["cvs-adc2", "cvs-adc2x", "cvs-adc3"])
def test_matvec_synthetic_cvs_shape(method):
    from adcc_b200.synthetic import restricted_scf_dict
    data = restricted_scf_dict(20, 5, seed=1004)
    check_matvec_parity(data, method, 1)


def test_davidson_synthetic_c1_adc2_3_singlets():
    # BASELINE configs[0]: ADC(2), 3 singlets
    from adcc_b200.synthetic import named_case
    data, _ = named_case("h2o_ccpvdz")
    check_run_parity(data, "adc2", "singlet", None, n=3, conv_tol=1e-6)   # run_adc default tolerance


def test_any_kind_unrestricted_style(h2o):
    check_run_parity(h2o, "adc2", "any", None, n=4)


@pytest.mark.parametrize("method", ["adc1", "adc2", "adc2x", "adc3", "cvs-adc2"])
def test_oscillator_strengths_vs_oracle(h2o, method):
    from adc_parity_common import check_oscillator_strengths
    check_oscillator_strengths(h2o, method, 1 if method.startswith("cvs") else None)


@pytest.mark.parametrize("case", ["gen", "cvs", "fc-fv-cvs"])
def test_import_blocks_device(case):
    from import_common import check_imports
    check_imports(case)


@pytest.mark.parametrize("case", ["gen", "fc-fv-cvs", "fv-cvs"])
def test_import_blocks_from_spatial_eri_device(case):
    from import_common import check_imports
    check_imports(case, spatial=True)


def test_shifted_and_projected_matrix_device(h2o):
    from adc_parity_common import check_shifted_projected
    check_shifted_projected(h2o)


def test_ao_basis_import_device(h2o):
    from adc_parity_common import check_ao_import
    check_ao_import(h2o, "adc2")
    from adcc_b200.synthetic import restricted_scf_dict
    check_ao_import(restricted_scf_dict(12, 4, seed=2), "adc2x")


@pytest.mark.parametrize("method", ["adc2", "adc3", "cvs-adc2x"])
def test_batched_matvec_device(h2o, method):
    from adc_parity_common import check_batched_matvec
    check_batched_matvec(h2o, method, 1 if method.startswith("cvs") else None)


def test_open_shell_unrestricted_reference_device():
    from adc_parity_common import check_open_shell
    check_open_shell()


@pytest.mark.parametrize("method", ["adc1", "adc2", "adc2x", "adc3", "cvs-adc2x"])
def test_graph_replay_device(h2o, method, monkeypatch):
    """CUDA-graph capture of the dense matvec: bit-identical to the eager launches, parity with the oracle."""
    from adc_parity_common import check_graph_replay
    check_graph_replay(h2o, method, 1 if method.startswith("cvs") else None, monkeypatch)


def test_extra_terms_plugin_interface_device(h2o):
    from adc_parity_common import check_extra_term
    check_extra_term(h2o)


@pytest.mark.parametrize("method,core,n_states", [("adc2", None, 7), ("adc3", None, 7), ("cvs-adc2x", 1, 7)])
def test_dense_export_device(h2o, method, core, n_states):
    """to_ndarray in the reference's dense basis: spectrum = converged states, and it contains the
    reference's in-source known answers (smoke_test.py:32-58)."""
    import numpy as np
    from adc_parity_common import check_dense_export
    from test_oracle_golden import SINGLETS, TRIPLETS
    spectrum = check_dense_export(h2o, method, core, n_states)
    for table in (SINGLETS, TRIPLETS):
        for val in table[method]:
            assert np.min(np.abs(spectrum - val)) < 1e-6, (method, val)


@pytest.mark.parametrize("method", ["adc2", "cvs-adc2x"])
def test_matvec_between_host_buffers_dense_store_device(h2o, method):
    import torch
    from adc_parity_common import build_pair, random_vector, rel_err, SIGMA_RTOL
    matrix, omatrix = build_pair(h2o, method, 1 if method.startswith("cvs") else None)
    v, ov = random_vector(matrix, omatrix, 5)
    u_host = {b: torch.from_numpy(ov[b].copy()).pin_memory() for b in ov}
    s_host = {b: torch.full_like(u_host[b], float("nan")).pin_memory() for b in ov}
    matrix.matvec_host(u_host, s_host)
    osig = omatrix.matvec(ov)
    for b in osig:
        assert rel_err(s_host[b].numpy(), osig[b]) < SIGMA_RTOL


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[1..3] at their exact spin-orbital shapes: restricted
# synthetic SCF data of the molecule's dimensions (real cc-pVXZ SCF data cannot be produced offline,
# SURVEY 8(d)); sigma + diagonal + per-block parity against the oracle as in the reference's
# adcc/tests/AdcMatrix_test.py:213-285.
# ------------------------------------------------------------------------------------------------
_ORACLE_CACHE = {}


def _oracle_matrix(case, method):
    """Oracle matrices of the larger shapes are expensive (ADC(3) at o=32/v=140: about a minute of
    host time); dense and packed product tests share one."""
    import oracle
    from adcc_b200.synthetic import named_case
    key = (case, method)
    if key not in _ORACLE_CACHE:
        data, core = named_case(case)
        oref = oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core or None)
        _ORACLE_CACHE[key] = (data, core or None, oracle.OracleAdcMatrix(method, oref))
    return _ORACLE_CACHE[key]


def test_c2_shape_adc2_sigma_and_oscillator_strengths():
    """configs[1]: H2O cc-pVTZ shape (o=10, v=106), ADC(2) + oscillator strengths."""
    import adcc_b200 as adcc
    from adc_parity_common import compare_matrices, check_oscillator_strengths
    data, core, om = _oracle_matrix("h2o_ccpvtz", "adc2")
    m = adcc.AdcMatrix("adc2", adcc.ReferenceState(data), doubles_storage="dense")
    assert m.axis_lengths == {"ph": 1060, "pphh": 1123600}
    compare_matrices(m, om)
    f = check_oscillator_strengths(data, "adc2", None, n=3, conv_tol=1e-7)
    assert f.shape == (3,)


@pytest.mark.parametrize("storage", ["dense", "packed"])
def test_c3_shape_adc3_sigma(storage):
    """configs[2]: methyloxirane cc-pVDZ shape (o=32, v=140), ADC(3) incl. the vvvv intermediates,
    on the dense store and on the antisymmetry-packed store."""
    import adcc_b200 as adcc
    from adc_parity_common import compare_matrices
    data, core, om = _oracle_matrix("methyloxirane_ccpvdz", "adc3")
    m = adcc.AdcMatrix("adc3", adcc.ReferenceState(data), doubles_storage=storage)
    assert m.axis_lengths == {"ph": 4480, "pphh": 20070400}
    if storage == "dense":
        compare_matrices(m, om)
    else:
        from packed_parity_common import _compare
        _compare(adcc, m, om)


@pytest.mark.parametrize("storage", ["dense", "packed"])
@pytest.mark.parametrize("case,n_ph,n_pphh", [("h2o_ccpvtz_cvs", 212, 179776), ("ne_ccpvtz_cvs", 100, 40000)])
def test_c4_shapes_cvs_adc2x_sigma(case, n_ph, n_pphh, storage):
    """configs[3]: H2O / Ne cc-pVTZ CVS-ADC(2)-x shapes (c=2, o=8, v=106 / 50), on the dense store and
    on the virtual-pair-packed store (adcc_b200/packed_cvs.py)."""
    import adcc_b200 as adcc
    from adc_parity_common import compare_matrices
    data, core, om = _oracle_matrix(case, "cvs-adc2x")
    m = adcc.AdcMatrix("cvs-adc2x", adcc.ReferenceState(data, core_orbitals=core), doubles_storage=storage)
    assert m.axis_lengths == {"ph": n_ph, "pphh": n_pphh}
    if storage == "dense":
        compare_matrices(m, om)
    else:
        from packed_parity_common import _compare_cvs
        _compare_cvs(adcc, m, om, against_dense=False)
