"""Second offline fixture of the reference (examples/0_basics/h2o_sto3g_hfdata.hdf5, PySCF RHF):
the repository's own HDF5 reader, the oracle and the product path on it."""
import os

import numpy as np
import pytest

import cpu_shim
import oracle
from oracle.hfdata import load_h2o_sto3g, load_h2o_sto3g_pyscf
from test_oracle_golden import SINGLETS, TRIPLETS

REF_FILE = "/root/reference/examples/0_basics/h2o_sto3g_hfdata.hdf5"


@pytest.fixture(autouse=True)
def shim(monkeypatch):
    cpu_shim.install(monkeypatch)


@pytest.fixture(scope="module")
def pyscf_h2o():
    return load_h2o_sto3g_pyscf()


def _flat(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference tree only exists in the build container")
def test_hdf5_reader_on_reference_file(pyscf_h2o):
    """Every dataset of the reference's HDF5 file (deflate-compressed chunks, scalars, booleans,
    variable-length strings, nested group) read by adcc_b200.hdf5io equals the committed fixture."""
    from adcc_b200 import hdf5io
    from adcc_b200.DataHfProvider import DataHfProvider, import_scf_results
    got, want = _flat(hdf5io.load(REF_FILE)), _flat(pyscf_h2o)
    assert sorted(got) == sorted(want)
    for key in want:
        if isinstance(want[key], str):
            assert got[key] == want[key]
        else:
            np.testing.assert_array_equal(np.asarray(got[key]), np.asarray(want[key]), err_msg=key)
    assert got["restricted"] is True or got["restricted"] == np.True_
    provider = import_scf_results(REF_FILE)                     # path -> provider, as in the reference
    assert isinstance(provider, DataHfProvider) and provider.n_orbs_alpha == 7 and provider.restricted
    with pytest.raises(ValueError):
        import_scf_results("/nonexistent/file.hdf5")
    with pytest.raises(hdf5io.Hdf5FormatError):
        hdf5io._Reader(b"not an hdf5 file at all")


def test_fixture_is_consistent_with_the_first_one(pyscf_h2o):
    """Same molecule and geometry as examples/water/data.py: SCF energy and orbital energies agree."""
    first = load_h2o_sto3g()
    assert abs(float(pyscf_h2o["energy_scf"]) - float(first["energy_scf"])) < 1e-8
    np.testing.assert_allclose(np.sort(pyscf_h2o["orben_f"]), np.sort(first["orben_f"]), atol=1e-7)
    np.testing.assert_allclose(pyscf_h2o["multipoles"]["nuclear_1"], first["multipoles"]["nuclear_1"], atol=1e-12)


@pytest.mark.parametrize("method", ["adc1", "adc2", "adc2x", "adc3", "cvs-adc2", "cvs-adc2x"])
def test_oracle_known_answers_on_pyscf_fixture(pyscf_h2o, method):
    """The in-source known answers (smoke_test.py:32-58) from the independent PySCF-generated data
    (full 14^4 spin-orbital eri_ffff import path)."""
    core = 1 if method.startswith("cvs") else None
    for table, kw in ((SINGLETS, "n_singlets"), (TRIPLETS, "n_triplets")):
        ref = table[method]
        state = oracle.run_adc(pyscf_h2o, method, conv_tol=1e-8, core_orbitals=core, **{kw: len(ref)})
        assert state.converged
        np.testing.assert_allclose(state.eigenvalues, ref, atol=1e-6)


@pytest.mark.parametrize("method", ["adc2", "adc3", "cvs-adc2x"])
def test_product_known_answers_on_pyscf_fixture(pyscf_h2o, method):
    import adcc_b200 as adcc
    core = 1 if method.startswith("cvs") else None
    ref = SINGLETS[method]
    state = adcc.run_adc(pyscf_h2o, method=method, n_singlets=len(ref), conv_tol=1e-8, core_orbitals=core,
                         output=None)
    assert state.converged
    np.testing.assert_allclose(state.excitation_energy, ref, atol=1e-6)
    ostate = oracle.run_adc(pyscf_h2o, method, n_singlets=len(ref), conv_tol=1e-8, core_orbitals=core)
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=1e-8)
    assert state.n_iter == ostate.n_iter


def test_reference_cases_of_the_fixture(pyscf_h2o):
    """The fixture lists the subspace cases the reference tests run on it (gen, cvs, fc, fv, ...):
    each one builds and gives the oracle's lowest singlet."""
    import ast
    import adcc_b200 as adcc
    cases = ast.literal_eval(pyscf_h2o["reference_cases"])
    assert set(cases) == {"gen", "cvs", "fc", "fv", "fv-cvs", "fc-fv"}
    for name, kwargs in cases.items():
        method = "cvs-adc2" if "cvs" in name else "adc2"
        state = adcc.run_adc(pyscf_h2o, method=method, n_singlets=1, conv_tol=1e-8, output=None, **kwargs)
        ostate = oracle.run_adc(pyscf_h2o, method, n_singlets=1, conv_tol=1e-8, **kwargs)
        np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=1e-8, err_msg=name)
