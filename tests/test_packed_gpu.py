"""Packed doubles kernels + engine + device-side synthetic generator on the B200 (C ABI path)."""
import numpy as np
import pytest

from packed_parity_common import (check_packed_matrix, check_packed_run, check_packed_spin_kinds, check_spin_conserving_rows,
                                  check_synthetic)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    import torch
    from adcc_b200 import backend
    torch.cuda.set_device(0)
    return backend


@pytest.mark.parametrize("no,nv", [(4, 6), (6, 34), (10, 38), (2, 2)])
def test_pack_unpack_expand_fold(be, no, nv):
    rng = np.random.default_rng(no * nv)
    d = rng.standard_normal((no, no, nv, nv))
    d = d - d.transpose(1, 0, 2, 3)
    d = d - d.transpose(0, 1, 3, 2)
    U = be.packed_pack(be.from_numpy(d), no, nv)
    npo, npv = no * (no - 1) // 2, nv * (nv - 1) // 2
    assert tuple(U.shape) == (npo, npv)
    assert np.allclose(be.packed_unpack(U, no, nv).cpu().numpy(), d, atol=1e-15)
    X = be.packed_expand(0, U, no, nv).cpu().numpy().reshape(no, nv, no, nv)
    assert np.array_equal(X, d.transpose(0, 2, 1, 3))
    po = [(i, j) for j in range(no) for i in range(j)]
    pv = [(a, b) for b in range(nv) for a in range(b)]
    Ue = be.packed_expand(1, U, no, nv).cpu().numpy().reshape(no, no, npv)
    assert np.array_equal(Ue, np.stack([d[:, :, a, b] for a, b in pv], axis=-1))
    Uh = be.packed_expand(2, U, no, nv).cpu().numpy().reshape(nv, npo, nv)
    assert np.array_equal(Uh, np.stack([d[i, j] for i, j in po], axis=1))
    # folds
    C = rng.standard_normal((no, no, npv))
    S0 = rng.standard_normal((npo, npv))
    S = be.from_numpy(S0)
    be.packed_fold_occ(S, be.from_numpy(C), no, nv, alpha=0.5)
    ref = S0 + 0.5 * np.stack([C[i, j] - C[j, i] for i, j in po])
    assert np.allclose(S.cpu().numpy(), ref, atol=1e-14)
    C2 = rng.standard_normal((npo, nv, nv))
    S = be.from_numpy(S0)
    be.packed_fold_virt(S, nv, nv, be.from_numpy(C2), be.int64_tensor([r * nv * nv for r in range(npo)]), alpha=-0.5)
    ref = S0 - 0.5 * np.stack([C2[:, a, b] - C2[:, b, a] for a, b in pv], axis=1)
    assert np.allclose(S.cpu().numpy(), ref, atol=1e-14)
    # pair_select
    T = rng.standard_normal((3, nv, nv, 5))
    sel = be.pair_select(be.from_numpy(T), 1).cpu().numpy()
    assert np.array_equal(sel, np.stack([T[:, a, b, :] for a, b in pv], axis=1))


def test_synth_fill_matches_definition(be):
    """The CUDA hash generator is bit-identical to the NumPy definition of the inputs."""
    import cpu_shim

    class MP:                      # collect the numpy reference implementation without patching
        def __init__(self): self.fns = {}
        def setattr(self, obj, name, val): self.fns[name] = val
    mp = MP()
    cpu_shim.install(mp)
    import torch
    no, nv, seed, scale = 6, 10, 20261019, 0.02
    npo, npv = 15, 45
    shapes = {"vvvv_packed": (npv, npv), "oooo_packed": (npo, npo), "ovov_jbkc": (no * nv, no * nv),
              "ovvv_jabc": (no * npv, nv), "ovvv_ajbc": (nv, no * npv), "ooov_ijkb": (no, npo * nv),
              "ooov_ijak": (npo * nv, no)}
    for layout, shape in shapes.items():
        dev = be.synth_fill(be.empty(shape), layout, no, nv, seed, scale).cpu().numpy()
        ref = mp.fns["synth_fill"](torch.empty(shape, dtype=torch.float64), layout, no, nv, seed, scale).numpy()
        assert np.array_equal(dev, ref), layout
    from adcc_b200.synthetic import synthetic_eri_block
    for blk in ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv"):
        shape = [no if c == "o" else nv for c in blk]
        dev = be.synth_fill(be.empty(shape), "dense", no, nv, seed, scale, block=blk).cpu().numpy()
        assert np.array_equal(dev, synthetic_eri_block(seed, blk, no, nv, scale)), blk
    # row_begin chunking of the big operand
    part = be.synth_fill(be.empty((10, npv)), "vvvv_packed", no, nv, seed, scale, row_begin=7).cpu().numpy()
    full = be.synth_fill(be.empty((npv, npv)), "vvvv_packed", no, nv, seed, scale).cpu().numpy()
    assert np.array_equal(part, full[7:17])


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_engine_h2o(method):
    check_packed_matrix(method)
    check_packed_run(method)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_packed_batched_matvec(method):
    from packed_parity_common import check_packed_batched
    from adcc_b200.synthetic import restricted_scf_dict
    check_packed_batched(method)
    check_packed_batched(method, restricted_scf_dict(24, 5, seed=3), n=5)
    check_packed_batched(method, restricted_scf_dict(7, 2, seed=3), n=4, tiny_memory=True)


@pytest.mark.parametrize("method", ["cvs-adc2", "cvs-adc2x"])
def test_packed_cvs_engine(method):
    """CVS variants on the virtual-pair-packed store: sigma / blocks / diagonal vs the oracle, Davidson
    singlets, triplets and "any" with the oracle's iteration counts."""
    from packed_parity_common import check_packed_cvs_matrix, check_packed_cvs_run
    from adcc_b200.synthetic import restricted_scf_dict
    check_packed_cvs_matrix(method)
    check_packed_cvs_matrix(method, restricted_scf_dict(9, 4, seed=2), core=2)
    check_packed_cvs_matrix(method, restricted_scf_dict(24, 5, seed=4), core=1)      # nv = 38: TMA GEMM paths
    check_packed_cvs_run(method)


@pytest.mark.parametrize("method", ["adc2x", "adc3"])
def test_packed_singlets_triplets_h2o(method):
    check_packed_spin_kinds(method)


@pytest.mark.parametrize("method", ["adc2x", "adc3"])
def test_packed_engine_c1_shape(method):
    from adcc_b200.synthetic import named_case
    data, _ = named_case("h2o_ccpvdz")
    check_packed_matrix(method, data)


@pytest.mark.parametrize("no,nv", [(6, 10), (10, 38), (16, 48)])
def test_synthetic_workload_vs_oracle(no, nv):
    check_synthetic(no, nv, seed=5, run=(nv <= 38))


@pytest.mark.parametrize("no,nv", [(4, 6), (6, 10), (10, 38), (12, 44)])
def test_synthetic_adc3_vs_oracle(no, nv):
    """ADC(3) on the synthetic workload: every operand comes from the pair-packed builder
    (exchange-ordered slabs, pair matvecs); sigma, blocks, diagonal vs the dense oracle."""
    check_synthetic(no, nv, seed=5, method="adc3", run=(nv == 10))


def test_adc3_packed_builder_vs_dense_intermediates():
    from packed_parity_common import check_adc3_builder_vs_dense
    from adcc_b200.synthetic import named_case, restricted_scf_dict
    check_adc3_builder_vs_dense()
    check_adc3_builder_vs_dense(restricted_scf_dict(7, 2, seed=3))
    check_adc3_builder_vs_dense(named_case("h2o_ccpvdz")[0])


@pytest.mark.parametrize("nv,no", [(6, 4), (38, 10), (50, 7)])
def test_pair_matvec_and_exchange_slabs(be, nv, no):
    """adccb_pair_matvec / adccb_vvvv_exchange_slab / adccb_ovvv_exchange_slab against NumPy."""
    rng = np.random.default_rng(nv)
    npv = nv * (nv - 1) // 2
    pv = [(a, b) for b in range(nv) for a in range(b)]
    aa, bb = np.array([p[0] for p in pv]), np.array([p[1] for p in pv])
    V2 = rng.standard_normal((nv, no, npv))                    # [s, f, BC]
    dense = np.zeros((nv, no, nv, nv))
    dense[:, :, aa, bb] = V2
    dense[:, :, bb, aa] = -V2
    X = rng.standard_normal((no, nv))
    V2d = be.from_numpy(V2)
    # out[s,c] = sum_f sum_d A_(s,f)[c,d] X[f,d]
    got = be.pair_matvec(V2d, nv, n_out=nv, n_red=no, row_stride_out=no, row_stride_red=1,
                         X=be.from_numpy(X), x_by_out=False).cpu().numpy()
    assert np.allclose(got, np.einsum("sfcd,fd->sc", dense, X), atol=1e-12)
    # out[f,a] += sum_s sum_b A_(s,f)[a,b] R[s,b]
    R = rng.standard_normal((nv, nv))
    out0 = rng.standard_normal((no, nv))
    out = be.from_numpy(out0)
    be.pair_matvec(V2d, nv, n_out=no, n_red=nv, row_stride_out=1, row_stride_red=no, X=be.from_numpy(R),
                   x_by_out=False, out=out, alpha=0.5, beta=1.0)
    assert np.allclose(out.cpu().numpy(), out0 + 0.5 * np.einsum("sfab,sb->fa", dense, R), atol=1e-12)
    # ovvv exchange slab: V2[c,(j,AD)] = <jc||ad> -> [al,c,j,d]
    a0, na = 1, min(3, nv - 1)
    got = be.ovvv_exchange_slab(V2d, no, nv, a0, na).cpu().numpy()
    assert np.array_equal(got, dense.transpose(2, 0, 1, 3)[a0:a0 + na])
    # vvvv exchange slab from the pair-packed block
    W = rng.standard_normal((npv, npv))
    full = np.zeros((nv, nv, nv, nv))
    for (P, (a, b)) in enumerate(pv):
        blk = np.zeros((nv, nv))
        blk[aa, bb] = W[P]
        blk[bb, aa] = -W[P]
        full[a, b], full[b, a] = blk, -blk
    got = be.vvvv_exchange_slab(be.from_numpy(W), nv, a0, na).cpu().numpy()      # [al,b,c,d] = <ac||bd>
    assert np.array_equal(got, full.transpose(0, 2, 1, 3)[a0:a0 + na])



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
