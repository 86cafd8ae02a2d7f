"""Shared checks for the packed engine (run through the CPU shim here and through CUDA on the GPU box)."""
import numpy as np

import oracle
from oracle.hfdata import load_h2o_sto3g
from oracle.synthetic import OracleSyntheticRef
from adc_parity_common import random_vector, rel_err, SIGMA_RTOL, ENERGY_ATOL


def _packed_vector(adcc, v):
    from adcc_b200.packed import PackedDoubles
    return adcc.AmplitudeVector(ph=v.ph, pphh=PackedDoubles.from_dense(v.pphh))


def _compare(adcc, m, om, seed=3):
    v, ov = random_vector(m, om, seed)
    vp = _packed_vector(adcc, v)
    sig, osig = m.matvec(vp), om.matvec(ov)
    assert rel_err(sig.ph.to_ndarray(), osig["ph"]) < SIGMA_RTOL
    assert rel_err(sig.pphh.to_ndarray(), osig["pphh"]) < SIGMA_RTOL
    for blk in ("ph_ph", "ph_pphh", "pphh_ph", "pphh_pphh"):
        outb, inb = blk.split("_")
        res = m.block_apply(blk, vp[inb])
        ores = om.blocks[blk]({inb: ov[inb]})[outb]
        assert rel_err(res.to_ndarray(), ores) < SIGMA_RTOL, blk
    assert rel_err(m.diagonal().ph.to_ndarray(), om.diagonal()["ph"]) < SIGMA_RTOL
    dd, od = m.diagonal().pphh.to_ndarray(), om.diagonal()["pphh"]
    mask = np.ones_like(od, dtype=bool)
    n_o, n_v = od.shape[0], od.shape[2]
    mask[np.arange(n_o), np.arange(n_o)] = False
    mask[:, :, np.arange(n_v), np.arange(n_v)] = False
    assert rel_err(dd[mask], od[mask]) < SIGMA_RTOL
    # hermiticity on the packed store
    w, _ = random_vector(m, om, seed + 1)
    wp = _packed_vector(adcc, w)
    lhs, rhs = vp @ m.matvec(wp), sig @ wp
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs))
    return sig, osig


def check_packed_matrix(method, data=None):
    import adcc_b200 as adcc
    data = data if data is not None else load_h2o_sto3g()
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
    om = oracle.OracleAdcMatrix(method, oracle.OracleRefState(oracle.OracleHf(data)))
    return _compare(adcc, m, om)


def check_packed_run(method, n=4):
    import adcc_b200 as adcc
    data = load_h2o_sto3g()
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
    st = adcc.run_adc(m, n_states=n, conv_tol=1e-9, output=None)
    ost = oracle.run_adc(data, method, n_states=n, conv_tol=1e-9)
    assert st.converged
    np.testing.assert_allclose(st.excitation_energy, ost.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    assert st.n_iter == ost.n_iter


def check_packed_spin_kinds(method, n=3):
    """Restricted singlets / triplets on the packed store: the spin projector and the singlet
    purification act on the unpacked tensor (solver/explicit_symmetrisation.py)."""
    import adcc_b200 as adcc
    data = load_h2o_sto3g()
    for kw in (dict(n_singlets=n), dict(n_triplets=n)):
        m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
        st = adcc.run_adc(m, conv_tol=1e-8, output=None, **kw)
        ost = oracle.run_adc(data, method, conv_tol=1e-8, **kw)
        assert st.converged and st.n_iter == ost.n_iter
        np.testing.assert_allclose(st.excitation_energy, ost.eigenvalues, atol=ENERGY_ATOL, rtol=0)


def check_spin_conserving_rows(method, data):
    """The spin-blocked engine under AdcMatrix.spin_conserving_vectors(): for vectors that are zero in
    the M_s-changing spin blocks (the guesses of every spin kind and what Davidson makes of them) the
    row-restricted class GEMMs give the same sigma as the all-rows path; outside the block nothing is
    assumed (random vectors are covered by check_packed_matrix)."""
    import adcc_b200 as adcc
    from adcc_b200.guess import guesses_any, guesses_singlet, guesses_triplet
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
    eng = m._engine
    assert eng._spin is not None and eng.ms_conserving is False
    vecs = []
    for make in (guesses_any, guesses_singlet, guesses_triplet):
        g = make(m, 3, block="ph") + make(m, 2, block="pphh")
        vecs += g
        vecs.append(m @ g[0] + 0.3 * (m @ g[-1]))          # a Krylov-type vector: still M_s-conserving
    for v in vecs:
        full = m @ v
        with m.spin_conserving_vectors():
            assert eng.ms_conserving is True
            fast = m @ v
        assert eng.ms_conserving is False
        for b in ("ph", "pphh"):
            ref = full[b]._data
            assert float((fast[b]._data - ref).abs().max()) <= 1e-13 * max(float(ref.abs().max()), 1e-300)
    flops_all = eng.flops["pphh_pphh/vvvv"]
    assert flops_all > 0


def check_packed_batched(method, data=None, n=3, tiny_memory=False):
    """AdcMatrix @ [vectors] on the packed store (PackedAdcEngine.matvec_batch: the big operands are
    streamed once per batch) equals the per-vector matvec and the oracle; Davidson on top of it
    keeps the oracle's iteration counts.  tiny_memory: force the batch to be cut into chunks."""
    import adcc_b200 as adcc
    from adcc_b200 import backend as be
    data = data if data is not None else load_h2o_sto3g()
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data), doubles_storage="packed")
    om = oracle.OracleAdcMatrix(method, oracle.OracleRefState(oracle.OracleHf(data)))
    vecs = [random_vector(m, om, 20 + k) for k in range(n)]
    packed = [_packed_vector(adcc, v) for v, _ in vecs]
    keep = be.free_memory
    if tiny_memory:
        be.free_memory = lambda: 1
    try:
        batched = m @ packed
    finally:
        be.free_memory = keep
    assert len(batched) == n
    for (v, ov), vp, res in zip(vecs, packed, batched):
        osig = om.matvec(ov)
        single = m.matvec(vp)
        for k in osig:
            assert rel_err(res[k].to_ndarray(), osig[k]) < SIGMA_RTOL, k
            assert rel_err(res[k].to_ndarray(), single[k].to_ndarray()) < 1e-13, k
    assert m @ [] == []
    return m


def check_packed_cvs_matrix(method, data=None, core=1, seed=11):
    """CVS-ADC(2) / CVS-ADC(2)-x on the virtual-pair-packed store (adcc_b200/packed_cvs.py) against
    the oracle: sigma, every block, both diagonals, hermiticity, dense round trip of the vector."""
    import adcc_b200 as adcc
    from adcc_b200.packed_cvs import CvsPackedDoubles, PackedCvsEngine
    data = data if data is not None else load_h2o_sto3g()
    m = adcc.AdcMatrix(method, adcc.ReferenceState(data, core_orbitals=core), doubles_storage="packed")
    assert isinstance(m._engine, PackedCvsEngine)
    om = oracle.OracleAdcMatrix(method, oracle.OracleRefState(oracle.OracleHf(data), core_orbitals=core))
    _compare_cvs(adcc, m, om, seed)
    return m, om


def _compare_cvs(adcc, m, om, seed=11, against_dense=True):
    from adcc_b200.packed_cvs import CvsPackedDoubles
    method = m.method.name
    v, ov = random_vector(m, om, seed)
    vp = adcc.AmplitudeVector(ph=v.ph, pphh=CvsPackedDoubles.from_dense(v.pphh))
    assert rel_err(vp.pphh.to_ndarray(), ov["pphh"]) < 1e-14
    assert abs(vp @ vp - v @ v) < 1e-11 * abs(v @ v)                       # every stored element counts twice
    sig, osig = m.matvec(vp), om.matvec(ov)
    assert isinstance(sig.pphh, CvsPackedDoubles)
    assert rel_err(sig.ph.to_ndarray(), osig["ph"]) < SIGMA_RTOL
    assert rel_err(sig.pphh.to_ndarray(), osig["pphh"]) < SIGMA_RTOL
    for blk in ("ph_ph", "ph_pphh", "pphh_ph", "pphh_pphh"):
        outb, inb = blk.split("_")
        res = m.block_apply(blk, vp[inb])
        ores = om.blocks[blk]({inb: ov[inb]})[outb]
        assert rel_err(res.to_ndarray(), ores) < SIGMA_RTOL, blk
    assert rel_err(m.diagonal().ph.to_ndarray(), om.diagonal()["ph"]) < SIGMA_RTOL
    dd, od = m.diagonal().pphh.to_ndarray(), om.diagonal()["pphh"]
    mask = np.ones_like(od, dtype=bool)
    n_v = od.shape[2]
    mask[:, :, np.arange(n_v), np.arange(n_v)] = False                     # a == b is not stored
    assert rel_err(dd[mask], od[mask]) < SIGMA_RTOL
    w, _ = random_vector(m, om, seed + 1)
    wp = adcc.AmplitudeVector(ph=w.ph, pphh=CvsPackedDoubles.from_dense(w.pphh))
    lhs, rhs = vp @ m.matvec(wp), sig @ wp
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs))
    if against_dense:      # the dense table interpreter gives the same sigma from the same (unpacked) vector
        md = adcc.AdcMatrix(method, m.reference_state, doubles_storage="dense")
        sd = md.matvec(v)
        assert rel_err(sig.pphh.to_ndarray(), sd.pphh.to_ndarray()) < SIGMA_RTOL


def check_packed_cvs_run(method, data=None, core=1, n=2, kinds=("singlet", "triplet", "any")):
    """Davidson on the packed CVS store: energies and iteration counts of the oracle's solver."""
    import adcc_b200 as adcc
    data = data if data is not None else load_h2o_sto3g()
    for kind in kinds:
        kw = {"singlet": {"n_singlets": n}, "triplet": {"n_triplets": n}, "any": {"n_states": n}}[kind]
        m = adcc.AdcMatrix(method, adcc.ReferenceState(data, core_orbitals=core), doubles_storage="packed")
        st = adcc.run_adc(m, conv_tol=1e-8, output=None, **kw)
        ost = oracle.run_adc(data, method, conv_tol=1e-8, core_orbitals=core, kind=kind, **kw)
        assert st.converged and st.n_iter == ost.n_iter, (kind, st.n_iter, ost.n_iter)
        np.testing.assert_allclose(st.excitation_energy, ost.eigenvalues, atol=ENERGY_ATOL, rtol=0)
    return st


def check_synthetic(no, nv, seed=5, run=True, method="adc2x"):
    import adcc_b200 as adcc
    from adcc_b200.packed import synthetic_reference_state
    ref = synthetic_reference_state(no, nv, seed=seed, scale=0.02)
    om = oracle.OracleAdcMatrix(method, OracleSyntheticRef(no, nv, seed=seed, scale=0.02))
    m = adcc.AdcMatrix(method, ref, doubles_storage="packed")
    _compare(adcc, m, om)
    if run:
        st = adcc.run_adc(m, n_states=3, conv_tol=1e-8, output=None)
        ost = oracle.run_adc(om, method, n_states=3, conv_tol=1e-8)
        assert st.converged
        np.testing.assert_allclose(st.excitation_energy, ost.eigenvalues, atol=ENERGY_ATOL, rtol=0)


def check_adc3_builder_vs_dense(data=None):
    """The ADC(3) operands built from pair-packed ERIs (adcc_b200/packed_adc3.py: exchange-ordered
    slabs, pair matvecs, packed GEMMs) against the dense intermediates of adcc_b200/adc_pp.py
    (adc3_m11 / adc3_pia / adc3_pib evaluated with plain einsum on dense blocks)."""
    import adcc_b200 as adcc
    from adcc_b200 import backend as be
    data = data if data is not None else load_h2o_sto3g()
    ref = adcc.ReferenceState(data)
    m = adcc.AdcMatrix("adc3", ref, doubles_storage="packed")
    eng = m._engine
    no, nv, npo, npv = eng.no, eng.nv, eng.npo, eng.npv
    dense = adcc.AdcMatrix("adc3", adcc.ReferenceState(data), doubles_storage="dense")
    im = dense.intermediates
    m11 = im.adc3_m11.to_ndarray().reshape(no * nv, no * nv)
    assert rel_err(eng.M1.cpu().numpy(), m11) < 1e-12
    pia, pib = im.adc3_pia.to_ndarray(), im.adc3_pib.to_ndarray()
    po = [(i, j) for j in range(no) for i in range(j)]
    pv = [(a, b) for b in range(nv) for a in range(b)]
    G2 = eng.G2.cpu().numpy().reshape(npo, nv, no)                       # [(IJ,a),k] = pia[ij,k,a]
    want = np.stack([pia[i, j].T for i, j in po])                        # [IJ, a, k]
    assert rel_err(G2, want) < 1e-12
    G1 = eng.G1.cpu().numpy().reshape(no, npo, nv)                       # [i,(JK,b)] = pia[jk,i,b]
    assert rel_err(G1, np.stack([pia[j, k] for j, k in po], axis=1)) < 1e-12
    V2 = eng.V2.cpu().numpy().reshape(nv, no, npv)                       # [a,(j,BC)] = pib[j,a,bc]
    want = np.stack([pib[:, :, bb, cc] for bb, cc in pv], axis=-1).transpose(1, 0, 2)
    assert rel_err(V2, want) < 1e-12
    V1 = eng.V1.contiguous().cpu().numpy().reshape(no, npv, nv)          # [(j,AB),c] = pib[j,c,ab] (view of V2)
    assert rel_err(V1, want.transpose(1, 2, 0)) < 1e-12
