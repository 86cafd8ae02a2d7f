"""Packed doubles store + ADC(2)/ADC(2)-x engine + synthetic workload: host logic against the
oracle through the TEST-ONLY CPU shim (see tests/cpu_shim.py)."""
import numpy as np
import pytest

import cpu_shim
from packed_parity_common import check_packed_matrix, check_packed_run, check_packed_spin_kinds, check_spin_conserving_rows, check_synthetic


@pytest.fixture(autouse=True)
def shim(monkeypatch):
    cpu_shim.install(monkeypatch)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_matvec_h2o(method):
    check_packed_matrix(method)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_davidson_h2o(method):
    check_packed_run(method)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_singlets_triplets_h2o(method):
    check_packed_spin_kinds(method)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_batched_matvec(method):
    """SURVEY 8(f2) on the packed store (adc3: the engine's per-vector path behind the same call)."""
    from packed_parity_common import check_packed_batched
    from adcc_b200.synthetic import restricted_scf_dict
    check_packed_batched(method)
    check_packed_batched(method, restricted_scf_dict(7, 2, seed=3), n=4, tiny_memory=True)


@pytest.mark.parametrize("method", ["cvs-adc2", "cvs-adc2x"])
def test_packed_cvs_matvec_and_davidson(method):
    """SURVEY 8(a) rows a8/a10/a11/a13/a14 (+CVS) on the virtual-pair-packed store."""
    from packed_parity_common import check_packed_cvs_matrix, check_packed_cvs_run
    from adcc_b200.synthetic import restricted_scf_dict
    check_packed_cvs_matrix(method)
    check_packed_cvs_matrix(method, restricted_scf_dict(9, 4, seed=2), core=2)       # two core orbitals
    check_packed_cvs_run(method)


def test_synthetic_workload_small():
    check_synthetic(6, 10, seed=5)


@pytest.mark.parametrize("no,nv", [(4, 6), (6, 10)])
def test_synthetic_adc3_small(no, nv):
    check_synthetic(no, nv, seed=5, method="adc3", run=(nv == 10))


def test_adc3_packed_builder_vs_dense_intermediates():
    from packed_parity_common import check_adc3_builder_vs_dense
    from adcc_b200.synthetic import restricted_scf_dict
    check_adc3_builder_vs_dense()
    check_adc3_builder_vs_dense(restricted_scf_dict(7, 2, seed=3))


def test_packed_tensor_semantics():
    import adcc_b200 as adcc
    from adcc_b200.packed import PackedDoubles
    from oracle.hfdata import load_h2o_sto3g
    ref = adcc.ReferenceState(load_h2o_sto3g())
    rng = np.random.default_rng(0)
    d = rng.standard_normal((10, 10, 4, 4))
    d = d - d.transpose(1, 0, 2, 3)
    d = d - d.transpose(0, 1, 3, 2)
    t = adcc.Tensor.from_ndarray(ref.mospaces, "o1o1v1v1", d)
    p = PackedDoubles.from_dense(t)
    assert p.shape == (10, 10, 4, 4) and p.packed_shape == (45, 6)
    np.testing.assert_allclose(p.to_ndarray(), d, atol=1e-15)
    assert abs(p.dot(p) - np.vdot(d, d)) < 1e-11          # dense Frobenius weight (4x)
    np.testing.assert_allclose((2.0 * p - p / 4.0).to_ndarray(), 1.75 * d, atol=1e-14)
    assert p.antisymmetrise(0, 1) is p and p.symmetrise([(0, 1), (2, 3)]) is p
    assert p[1, 0, 2, 3] == -p[0, 1, 2, 3] and p[0, 0, 1, 2] == 0.0
    with pytest.raises(ValueError):
        p + t                                              # mixed storage
    q = p.zeros_like()
    q[0, 1, 2, 3] = 2.0
    assert q[1, 0, 3, 2] == 2.0 and q[0, 1, 3, 2] == -2.0
    with pytest.raises(ValueError):
        PackedDoubles.zeros(ref.mospaces, ["o1", "o1", "v1", "v1"]).set_from_ndarray(np.ones((10, 10, 4, 4)))


def test_doubles_guesses_on_packed_stores():
    """Doubles guesses (adcc/guess/guesses_from_diagonal.py:233-305, fill_pp_doubles_guesses) written
    through element access into the packed stores equal the dense-store guesses."""
    import adcc_b200 as adcc
    from oracle.hfdata import load_h2o_sto3g
    data = load_h2o_sto3g()
    for core, method in ((None, "adc2x"), (1, "cvs-adc2x")):
        ref = adcc.ReferenceState(data, core_orbitals=core)
        packed = adcc.AdcMatrix(method, ref, doubles_storage="packed")
        dense = adcc.AdcMatrix(method, ref, doubles_storage="dense")
        for fn in (adcc.guesses_singlet, adcc.guesses_triplet, adcc.guesses_any):
            gp, gd = fn(packed, 4, block="pphh"), fn(dense, 4, block="pphh")
            assert len(gp) == len(gd) == 4
            for a, b in zip(gp, gd):
                np.testing.assert_allclose(a.pphh.to_ndarray(), b.pphh.to_ndarray(), atol=1e-15)
                assert abs(a @ a - 1.0) < 1e-13


def test_cvs_packed_tensor_semantics():
    import adcc_b200 as adcc
    from adcc_b200.packed_cvs import CvsPackedDoubles
    from oracle.hfdata import load_h2o_sto3g
    ref = adcc.ReferenceState(load_h2o_sto3g(), core_orbitals=1)
    rng = np.random.default_rng(0)
    d = rng.standard_normal((8, 2, 4, 4))
    d = d - d.transpose(0, 1, 3, 2)
    t = adcc.Tensor.from_ndarray(ref.mospaces, "o1o2v1v1", d)
    p = CvsPackedDoubles.from_dense(t)
    assert p.shape == (8, 2, 4, 4) and p.packed_shape == (16, 6) and p.size == 256
    np.testing.assert_allclose(p.to_ndarray(), d, atol=1e-15)
    assert abs(p.dot(p) - np.vdot(d, d)) < 1e-11          # dense Frobenius weight (2x)
    np.testing.assert_allclose(p.dot([p, 2.0 * p]), [np.vdot(d, d), 2 * np.vdot(d, d)], rtol=1e-13)
    np.testing.assert_allclose((2.0 * p - p / 4.0).to_ndarray(), 1.75 * d, atol=1e-14)
    assert p.antisymmetrise(2, 3) is p
    assert p[3, 1, 2, 3] == -p[3, 1, 3, 2] == d[3, 1, 2, 3] and p[0, 0, 1, 1] == 0.0
    with pytest.raises(ValueError):
        p + t                                              # mixed storage
    q = p.zeros_like()
    q[5, 1, 3, 2] = 2.0
    assert q[5, 1, 2, 3] == -2.0 and type(q) is CvsPackedDoubles
    with pytest.raises(ValueError):
        CvsPackedDoubles.zeros(ref.mospaces, ["o1", "o2", "v1", "v1"]).set_from_ndarray(np.ones((8, 2, 4, 4)))
    # a non-antisymmetric dense tensor is antisymmetrised on the way in
    e = rng.standard_normal((8, 2, 4, 4))
    pe = CvsPackedDoubles.from_dense(adcc.Tensor.from_ndarray(ref.mospaces, "o1o2v1v1", e))
    np.testing.assert_allclose(pe.to_ndarray(), 0.5 * (e - e.transpose(0, 1, 3, 2)), atol=1e-15)
    with pytest.raises(NotImplementedError):
        adcc.AdcMatrix("cvs-adc3", ref, doubles_storage="packed")


def test_synthetic_definition_symmetries():
    from adcc_b200.synthetic import synthetic_eri_block
    for blk, bra, ket, swap in (("vvvv", True, True, True), ("oooo", True, True, True),
                                ("oovv", True, True, False), ("ooov", True, False, False),
                                ("ovvv", False, True, False), ("ovov", False, False, True)):
        v = synthetic_eri_block(11, blk, 4, 6, 0.02)
        if bra:
            assert np.array_equal(v, -v.transpose(1, 0, 2, 3))
        if ket:
            assert np.array_equal(v, -v.transpose(0, 1, 3, 2))
        if swap:
            assert np.array_equal(v, v.transpose(2, 3, 0, 1))
        assert 0.015 < v[v != 0].std() < 0.025


@pytest.mark.parametrize("method", ["adc2x", "adc3"])
def test_spin_blocked_operands_restricted_reference(method, monkeypatch):
    """Restricted references keep only the spin-conserving diagonal blocks of W / O / Y (forced on
    here; by default from 64 virtual spin orbitals up): same sigma, blocks, diagonal and states."""
    import adcc_b200 as adcc
    from oracle.hfdata import load_h2o_sto3g
    from adcc_b200.synthetic import restricted_scf_dict
    monkeypatch.setenv("ADCCB_SPIN_BLOCKS", "1")
    for data in (load_h2o_sto3g(), restricted_scf_dict(9, 3, seed=11)):
        m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
        assert m._engine._spin is not None and m._engine.W is None
        sp = m._engine._spin
        assert sp["W"][2] is sp["W"][0] and sp["O"][2] is sp["O"][0]        # beta-beta stored once
        check_packed_matrix(method, data)
        check_spin_conserving_rows(method, data)
    check_packed_run(method)                 # through run_adc: row-restricted class GEMMs
    check_packed_spin_kinds(method)
