"""Host logic (einsum planner, block tables, Davidson, guesses, workflow) against the oracle,
on CPU tensors through the TEST-ONLY shim tests/cpu_shim.py (the build container has no GPU).
The same checks run through the real CUDA backend in tests/test_adc_gpu.py."""
import numpy as np
import pytest

import oracle
from oracle.hfdata import load_h2o_sto3g

import cpu_shim
from adc_parity_common import (METHODS, check_matvec_parity, check_run_parity, SINGLETS)


@pytest.fixture(autouse=True)
def shim(monkeypatch):
    cpu_shim.install(monkeypatch)


@pytest.fixture(scope="module")
def h2o():
    return load_h2o_sto3g()


@pytest.mark.parametrize("method", METHODS)
def test_matvec_vs_oracle(h2o, method):
    check_matvec_parity(h2o, method, 1 if method.startswith("cvs") else None)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("kind", ["singlet", "triplet", "any"])
def test_run_adc_vs_oracle(h2o, method, kind):
    check_run_parity(h2o, method, kind, 1 if method.startswith("cvs") else None)


def test_synthetic_restricted_small():
    from adcc_b200.synthetic import restricted_scf_dict
    data = restricted_scf_dict(9, 3, seed=7)
    for method in ("adc2", "adc2x", "adc3"):
        check_matvec_parity(data, method, None)
    data = restricted_scf_dict(9, 4, seed=8)
    check_matvec_parity(data, "cvs-adc2x", 1)
    check_run_parity(data, "cvs-adc2", "singlet", 1, n=2)


def test_space_and_reference_properties(h2o):
    import adcc_b200 as adcc
    ref = adcc.ReferenceState(h2o, core_orbitals=1, frozen_virtual=1)
    ms = ref.mospaces
    assert ms.core_orbitals == [0, 7] and ms.occupied_orbitals == [1, 2, 3, 4, 8, 9, 10, 11]
    assert ms.virtual_orbitals == [5, 12] and ms.frozen_virtual == [6, 13] and ms.frozen_core == []
    assert ref.is_aufbau_occupation
    c, ca, cb = (f("o1b") for f in (ref.orbital_coefficients, ref.orbital_coefficients_alpha,
                                    ref.orbital_coefficients_beta))
    assert c.shape == (8, 7) and np.array_equal(ca + cb, c)
    assert not ca[4:].any() and not cb[:4].any() and np.array_equal(ca[:4], cb[4:])     # restricted
    with pytest.raises(ValueError):
        ref.orbital_coefficients("o1")


def test_api_surface_and_errors(h2o):
    import adcc_b200 as adcc
    with pytest.raises(adcc.InputError):
        adcc.run_adc(h2o, method="adc2")                       # no states requested
    with pytest.raises(adcc.InputError):
        adcc.run_adc(h2o, method="cvs-adc2", n_singlets=1)      # core_orbitals missing
    with pytest.raises(adcc.InputError):
        adcc.run_adc(h2o, method="adc7", n_singlets=1)
    with pytest.raises(adcc.InputError):
        adcc.run_adc(h2o, method="adc2", n_singlets=1, n_triplets=1)
    refstate = adcc.ReferenceState(h2o)
    with pytest.raises(ValueError):
        adcc.AdcMatrix("cvs-adc2", refstate)
    matrix = adcc.AdcMatrix("adc2", refstate)
    assert matrix.axis_blocks == ["ph", "pphh"]
    assert matrix.shape == (40 + 1600, 40 + 1600)
    assert matrix.block_orders == {"ph_ph": 2, "ph_pphh": 1, "pphh_ph": 1, "pphh_pphh": 0}
    guesses = adcc.guesses_singlet(matrix, 3)
    with pytest.raises(ValueError):
        adcc.jacobi_davidson(matrix, guesses, n_ep=5)
    with pytest.raises(TypeError):
        adcc.jacobi_davidson("nomatrix", guesses)
    # n_applies bookkeeping (adcc/tests/solver/davidson_test.py:60-62)
    state = adcc.jacobi_davidson(matrix, guesses, n_ep=2, conv_tol=1e-6)
    assert state.n_applies >= len(guesses)
    assert state.converged


def test_lincomb_many(h2o):
    """lincomb_many = several lincomb calls on the same inputs (blocks of AmplitudeVectors included)."""
    import adcc_b200 as adcc
    from adcc_b200.functions import lincomb, lincomb_many
    m = adcc.AdcMatrix("adc2", adcc.ReferenceState(h2o))
    vecs = [adcc.guess_zero(m).set_random() for _ in range(4)]
    C = np.random.default_rng(1).standard_normal((3, 4))
    outs = lincomb_many(C, vecs)
    assert len(outs) == 3
    for row, o in zip(C, outs):
        want = lincomb(row, vecs, evaluate=True)
        for b in want.blocks:
            np.testing.assert_allclose(o[b].to_ndarray(), want[b].to_ndarray(), rtol=0, atol=1e-14)
    assert lincomb_many(np.zeros((0, 4)), vecs) == []
    with pytest.raises(ValueError):
        lincomb_many(np.ones((2, 3)), vecs)


def test_tensor_algebra_semantics(h2o):
    """Closed-form expectations in the style of libadcc_src/tests/TensorTest.cc:99-970 and
    adcc/tests/einsum_test.py (numpy.einsum as the reference, rtol 1e-10 / atol 1e-14)."""
    import adcc_b200 as adcc
    from adcc_b200.functions import einsum, direct_sum, tensordot
    ref = adcc.ReferenceState(h2o)
    rng = np.random.default_rng(0)
    oovv = ref.oovv
    ovov = ref.ovov
    a = oovv.to_ndarray()
    bnd = ovov.to_ndarray()
    np.testing.assert_allclose(einsum("ijab,kbjc->ikac", oovv, ovov).to_ndarray(),
                               np.einsum("ijab,kbjc->ikac", a, bnd), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(einsum("ijab,ijab->", oovv, oovv), np.einsum("ijab,ijab->", a, a))
    np.testing.assert_allclose(einsum("iaia->ia", ovov).to_ndarray(), np.einsum("iaia->ia", bnd))
    np.testing.assert_allclose(einsum("ijab->baji", oovv).to_ndarray(), a.transpose(3, 2, 1, 0))
    np.testing.assert_allclose(tensordot(oovv, ovov, axes=([2, 3], [1, 3])).to_ndarray(),
                               np.tensordot(a, bnd, axes=([2, 3], [1, 3])), rtol=1e-10, atol=1e-14)
    # symmetrise / antisymmetrise include the 1/2 (Tensor.hh:229-244)
    np.testing.assert_allclose(oovv.antisymmetrise(0, 1).to_ndarray(), 0.5 * (a - a.transpose(1, 0, 2, 3)))
    np.testing.assert_allclose(oovv.symmetrise((0, 1), (2, 3)).to_ndarray(),
                               0.5 * (a + a.transpose(1, 0, 3, 2)))
    with pytest.raises(ValueError):
        ovov.symmetrise(0, 1)                      # different subspaces
    # direct_sum and diagonal conventions (Appendix A.5)
    fo, fv = ref.foo.diagonal(), ref.fvv.diagonal()
    ds = direct_sum("a-i->ia", fv, fo).to_ndarray()
    np.testing.assert_allclose(ds, -fo.to_ndarray()[:, None] + fv.to_ndarray()[None, :])
    assert ovov.diagonal(0, 2).shape == (4, 4, 10)
    # scalar +/- 0 returns self, arithmetic
    assert (oovv + 0) is oovv
    np.testing.assert_allclose((2.0 * oovv - oovv / 2.0).to_ndarray(), 1.5 * a)
    np.testing.assert_allclose((oovv * oovv).to_ndarray(), a * a)
    assert abs(oovv.dot(oovv) - np.vdot(a, a)) < 1e-12
    np.testing.assert_allclose(oovv.dot([oovv, oovv]), [np.vdot(a, a)] * 2)
    # element access honours symmetry
    sym = adcc.Symmetry(ref.mospaces, "o1o1v1v1", permutations=["ijab", "-jiab", "-ijba"])
    t = adcc.Tensor(sym)
    t[0, 1, 2, 3] = 1.0
    assert t[1, 0, 2, 3] == -1.0 and t[1, 0, 3, 2] == 1.0
    assert not t.is_allowed((1, 1, 2, 3))
    with pytest.raises(RuntimeError):
        t[1, 1, 2, 3] = 1.0
    assert t.select_n_max(1)[0][0] == (0, 1, 2, 3)
    # immutability of imported tensors (ReferenceState.cc:471,507)
    with pytest.raises(RuntimeError):
        oovv.fill(1.0)


@pytest.mark.parametrize("method", ["adc1", "adc2", "adc2x", "adc3", "cvs-adc2"])
def test_oscillator_strengths_vs_oracle(h2o, method):
    from adc_parity_common import check_oscillator_strengths
    check_oscillator_strengths(h2o, method, 1 if method.startswith("cvs") else None)


def test_shifted_and_projected_matrix(h2o):
    """AdcMatrixShifted / AdcMatrixProjected (adcc/AdcMatrix.py:649-780) against the
    oracle matvec plus NumPy masks; core excitations by projection converge to the CVS energies'
    neighbourhood (examples/2_core_excitations/water_core_by_projection.py)."""
    from adc_parity_common import check_shifted_projected
    check_shifted_projected(h2o)


def test_ao_basis_import(h2o):
    from adc_parity_common import check_ao_import
    check_ao_import(h2o, "adc2")


def test_unrestricted_style_kinds():
    """Unrestricted-style reference (no spin-block maps): kind='any' and kind='spin_flip' through the
    dense engine against the oracle (adcc/workflow.py:288-348, guess kwargs adcc/guess/__init__.py:30-45)."""
    from adc_parity_common import check_unrestricted_kinds
    check_unrestricted_kinds()


@pytest.mark.parametrize("method", ["adc1", "adc2x", "adc3", "cvs-adc3"])
def test_batched_matvec(h2o, method):
    from adc_parity_common import check_batched_matvec
    check_batched_matvec(h2o, method, 1 if method.startswith("cvs") else None)


def test_open_shell_unrestricted_reference():
    """n_alpha != n_beta (ragged alpha/beta halves on every axis), <pq||rs> import path."""
    from adc_parity_common import check_open_shell
    check_open_shell()


def test_tensor_symmetrise_reference_goldens():
    from tensor_golden_common import check_tensor_goldens
    check_tensor_goldens()


def test_tensor_ops_closed_form(h2o):
    """More of the closed-form expectations of libadcc_src/tests/TensorTest.cc:99-970: trace, set_mask,
    select_n_*, matmul, direct_sum of three operands, copy / *_like, in-place ops, random fill."""
    import adcc_b200 as adcc
    from adcc_b200.functions import einsum, direct_sum, trace
    ref = adcc.ReferenceState(h2o)
    ms = ref.mospaces
    rng = np.random.default_rng(1)
    a = rng.standard_normal((10, 10))
    A = adcc.Tensor.from_ndarray(ms, "o1o1", a)
    B = adcc.Tensor.from_ndarray(ms, "o1o1", a.T.copy())
    # trace / full contractions
    assert abs(trace("ii", A) - np.trace(a)) < 1e-13
    oooo = ref.oooo
    assert abs(trace("ijij", oooo) - np.einsum("ijij->", oooo.to_ndarray())) < 1e-12
    # matmul contracts axis 1 of a with axis 0 of b (export_Tensor.cc:420-422)
    np.testing.assert_allclose((A @ B).to_ndarray(), a @ a.T, atol=1e-13)
    np.testing.assert_allclose(A.T.to_ndarray(), a.T)
    # set_mask: Kronecker delta and general masks
    d = A.zeros_like()
    d.set_mask("ii", 1.0)
    np.testing.assert_array_equal(d.to_ndarray(), np.eye(10))
    t4 = ref.oovv.zeros_like()
    t4.set_mask("ijab", 2.0)
    assert np.all(t4.to_ndarray() == 2.0)
    t4.set_mask("iiab", 7.0)
    arr = t4.to_ndarray()
    assert np.all(arr[np.arange(10), np.arange(10)] == 7.0) and arr[0, 1, 0, 0] == 2.0
    # select_n_* (TensorImpl.cc:260-282): sorted, index + value
    sel = A.select_n_min(3)
    flat = np.sort(a, axis=None)
    assert [v for _, v in sel] == list(flat[:3])
    assert all(a[idx] == v for idx, v in sel)
    assert A.select_n_absmax(1)[0][1] == a.flat[np.argmax(np.abs(a))]
    # direct_sum with three operands and a permutation
    fo, fv = ref.foo.diagonal(), ref.fvv.diagonal()
    ds = direct_sum("-i+a-j->jai", fo, fv, fo).to_ndarray()
    o, v = fo.to_ndarray(), fv.to_ndarray()
    np.testing.assert_allclose(ds, (-o[:, None, None] + v[None, :, None] - o[None, None, :]).transpose(2, 1, 0))
    # copies, *_like, in-place arithmetic
    C = A.copy()
    C *= 2.0
    C += A
    C -= B
    C /= 4.0
    np.testing.assert_allclose(C.to_ndarray(), (3.0 * a - a.T) / 4.0, atol=1e-15)
    # augmented assignment rebinds, it does not touch the object named before (export_Tensor.cc:352-410:
    # "return self = self->add(other)"); the reference's state_diffdm.py:70-86 depends on it
    D = A.copy()
    alias = D
    D += A
    D *= 3.0
    np.testing.assert_allclose(alias.to_ndarray(), a, atol=0)
    np.testing.assert_allclose(D.to_ndarray(), 6.0 * a, atol=1e-15)
    frozen = A.copy()
    frozen.set_immutable()
    frozen += A                                  # fine: a new tensor
    np.testing.assert_allclose(frozen.to_ndarray(), 2.0 * a, atol=1e-15)
    assert np.all(A.ones_like().to_ndarray() == 1.0) and np.all(A.empty_like().to_ndarray() == 0.0)
    assert A.nosym_like().symmetry is None
    r = adcc.Tensor(adcc.Symmetry(ms, "o1o1v1v1", permutations=["ijab", "-jiab", "-ijba"])).set_random()
    rr = r.to_ndarray()
    assert np.allclose(rr, -rr.transpose(1, 0, 2, 3)) and np.allclose(rr, -rr.transpose(0, 1, 3, 2))
    # partial traces are not supported (opt_einsum_integration.py:68-71); extent mismatch is an error
    with pytest.raises(NotImplementedError):
        einsum("iiab->ab", ref.oovv)
    with pytest.raises(ValueError):
        einsum("ij,jk->ik", A, ref.fvv)


@pytest.mark.parametrize("method", ["adc2", "cvs-adc2x"])
def test_graph_replay_host_logic(h2o, method, monkeypatch):
    """Static-buffer / replay / copy-out protocol of the captured matvec (the shim re-runs eagerly)."""
    from adc_parity_common import check_graph_replay
    monkeypatch.setenv("ADCCB_SHIM_GRAPHS", "1")
    check_graph_replay(h2o, method, 1 if method.startswith("cvs") else None, monkeypatch)


@pytest.mark.parametrize("method,core,fc,fv", [("adc2", None, 1, None), ("adc2x", None, None, 1),
                                               ("adc3", None, 1, 1), ("cvs-adc2x", 1, 1, None),
                                               ("cvs-adc2", 1, None, 1), ("cvs-adc3", 1, 1, 1)])
def test_frozen_spaces_vs_oracle(h2o, method, core, fc, fv):
    from adc_parity_common import check_frozen_spaces
    check_frozen_spaces(h2o, method, core, fc, fv)


def test_einsum_planner_random_expressions(h2o):
    """Randomly generated 2- and 3-operand contractions over occupied / virtual axes (with
    transposed views as inputs, outer products, full contractions) against numpy.einsum; the
    reference's own einsum tests use the same oracle and tolerance (adcc/tests/einsum_test.py:36)."""
    import adcc_b200 as adcc
    from adcc_b200.functions import einsum
    ref = adcc.ReferenceState(h2o)
    rng = np.random.default_rng(2024)
    letters = {"o1": "ijklmn", "v1": "abcdef"}

    def rand_tensor(spaces):
        shape = [ref.mospaces.n_orbs(s) for s in spaces]
        arr = rng.standard_normal(shape)
        return adcc.Tensor.from_ndarray(ref.mospaces, spaces, arr), arr

    n_done = 0
    while n_done < 40:
        n_ops = int(rng.integers(2, 4))
        pool = {sp: list(letters[sp][:int(rng.integers(2, 5))]) for sp in letters}
        space_of = {c: sp for sp, cs in pool.items() for c in cs}
        ops, arrays, subs = [], [], []
        for _ in range(n_ops):
            nd = int(rng.integers(1, 5))
            idx = list(rng.choice(sorted(space_of), size=nd, replace=False))
            t, a = rand_tensor([space_of[c] for c in idx])
            if nd > 1 and rng.random() < 0.4:               # hand over a transposed (strided) view
                perm = [int(p) for p in rng.permutation(nd)]
                t, a, idx = t.transpose(perm), a.transpose(perm), [idx[p] for p in perm]
            ops.append(t), arrays.append(a), subs.append("".join(idx))
        flat = "".join(subs)
        counts = {c: flat.count(c) for c in set(flat)}
        if any(v > 2 for v in counts.values()):
            continue                                        # an index on three operands: not a pair contraction
        free = [c for c in dict.fromkeys(flat) if counts[c] == 1]
        out = "".join(rng.permutation(free)) if free else ""
        expr = ",".join(subs) + "->" + out
        expected = np.einsum(expr, *arrays)
        got = einsum(expr, *ops)
        got = got.to_ndarray() if hasattr(got, "to_ndarray") else got
        np.testing.assert_allclose(got, expected, rtol=1e-10, atol=1e-12, err_msg=expr)
        n_done += 1


def test_extra_terms_plugin_interface(h2o):
    from adc_parity_common import check_extra_term
    check_extra_term(h2o)


@pytest.mark.parametrize("method,core,n_states", [("adc1", None, 5), ("adc2", None, 7), ("adc2x", None, 7),
                                                  ("adc3", None, 7), ("cvs-adc1", 1, 2),
                                                  ("cvs-adc2", 1, 3), ("cvs-adc2x", 1, 7)])
def test_dense_export(h2o, method, core, n_states):
    from adc_parity_common import check_dense_export
    spectrum = check_dense_export(h2o, method, core, n_states)
    # the dense spectrum also contains the reference's in-source known answers (smoke_test.py:32-58)
    from test_oracle_golden import SINGLETS, TRIPLETS
    for table in (SINGLETS, TRIPLETS):
        for val in table[method]:
            assert np.min(np.abs(spectrum - val)) < 1e-6, (method, val)


@pytest.mark.parametrize("kind", ["singlet", "triplet"])
def test_transfer_cvs_to_full(h2o, kind):
    from adc_parity_common import check_transfer_cvs_to_full
    check_transfer_cvs_to_full(h2o, kind)


def test_cis_is_adc1_with_zeroth_order_properties(h2o):
    """adcc.cis (adcc/__init__.py:93-96): ADC(1) energies, ISR(0) transition densities."""
    import adcc_b200 as adcc
    import oracle
    from oracle.properties import oscillator_strengths
    state = adcc.cis(h2o, n_singlets=3, conv_tol=1e-9, output=None)
    ostate = oracle.run_adc(h2o, "adc1", n_singlets=3, conv_tol=1e-9)
    omatrix = oracle.OracleAdcMatrix("adc1", oracle.OracleRefState(oracle.OracleHf(h2o)))
    np.testing.assert_allclose(state.excitation_energy, ostate.eigenvalues, atol=1e-8)
    assert state.property_method == "adc0" and state.method.name == "adc1"
    f0, _ = oscillator_strengths(omatrix, ostate, level=0)
    f1, _ = oscillator_strengths(omatrix, ostate)
    np.testing.assert_allclose(state.oscillator_strength, f0, rtol=1e-6, atol=1e-10)
    assert np.max(np.abs(f0 - f1)) > 1e-4                  # the two levels really differ
    assert "adcc_b200" in adcc.banner()
    assert state.kind == "singlet"
    # describe(): the reference's table layout (adcc/ExcitedStates.py:50-166)
    lines = state.describe().splitlines()
    assert len({len(ln) for ln in lines}) == 1                      # a proper frame
    assert lines[1].startswith("| adc1 (isr0)") and lines[1].rstrip(" |").endswith("singlet, converged")
    assert "excitation energy" in lines[3] and "osc str" in lines[3] and "|v1|^2" in lines[3]
    assert "|v2|^2" not in lines[3] and "(au)" in lines[4] and len(lines) == 5 + 3 + 1
    assert "transition dipole moment" in state.describe(transition_dipole_moments=True)
    assert "osc str" not in state.describe(oscillator_strengths=False)


def test_chained_methods_reuse_ground_state_and_vectors(h2o):
    """The overview notebook's flow (examples/0_basics/00_Overview.ipynb): ADC(2) vectors as guesses
    for ADC(2)-x on the same ground state, those for ADC(3); a ReferenceState as first argument."""
    import adcc_b200 as adcc
    from test_oracle_golden import SINGLETS
    state = adcc.adc2(h2o, n_singlets=3, conv_tol=1e-8, output=None)
    s2x = adcc.adc2x(state.ground_state, n_singlets=3, guesses=state.excitation_vector, conv_tol=1e-8, output=None)
    s3 = adcc.adc3(s2x.ground_state, n_singlets=3, guesses=s2x.excitation_vector, conv_tol=1e-8, output=None)
    assert s2x.ground_state is state.ground_state and s3.ground_state is state.ground_state
    np.testing.assert_allclose(s2x.excitation_energy, SINGLETS["adc2x"], atol=1e-6)
    np.testing.assert_allclose(s3.excitation_energy, SINGLETS["adc3"], atol=1e-6)
    s1 = adcc.adc1(state.reference_state, n_singlets=3, conv_tol=1e-8, output=None)
    np.testing.assert_allclose(s1.excitation_energy, SINGLETS["adc1"], atol=1e-6)
    with pytest.raises(adcc.exceptions.InputError):
        adcc.adc2x(state.ground_state, n_singlets=3, guesses=state.excitation_vector[:2], output=None)


def test_davidson_edge_cases(h2o):
    from adc_parity_common import check_davidson_edge_cases
    check_davidson_edge_cases(h2o)


@pytest.mark.parametrize("method", ["adc2", "cvs-adc2x"])
def test_matvec_between_host_buffers_dense_store(h2o, method):
    """AdcMatrix.matvec_host on the dense store: sigma lands in the caller's host buffers."""
    import torch
    from adc_parity_common import build_pair, random_vector, rel_err, SIGMA_RTOL
    matrix, omatrix = build_pair(h2o, method, 1 if method.startswith("cvs") else None)
    v, ov = random_vector(matrix, omatrix, 5)
    u_host = {b: torch.from_numpy(ov[b].copy()) for b in ov}
    s_host = {b: torch.full_like(u_host[b], float("nan")) for b in ov}
    matrix.matvec_host(u_host, s_host)
    osig = omatrix.matvec(ov)
    for b in osig:
        assert rel_err(s_host[b].numpy(), osig[b]) < SIGMA_RTOL


def test_lanczos_solver_api(h2o):
    """The second eigensolver behind run_adc (adcc/solver/lanczos.py:255-311): both ends of the
    spectrum against the dense export of the matrix, thick restarts, the true residuals, the
    reference's warnings and argument checks."""
    import warnings
    import scipy.linalg as la
    import adcc_b200 as adcc
    from adcc_b200.guess import guesses_any
    from adcc_b200.solver.lanczos import lanczos, LanczosIterator
    matrix = adcc.AdcMatrix("adc2", adcc.ReferenceState(h2o))
    spectrum = np.linalg.eigvalsh(matrix.to_ndarray())
    guesses = guesses_any(matrix, 3, block="ph")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        low = lanczos(matrix, guesses, n_ep=2, which="SA", conv_tol=1e-8, debug_checks=True)
        high = lanczos(matrix, guesses, n_ep=2, which="LA", conv_tol=1e-8)
    assert low.converged and high.converged and low.algorithm == "lanczos"
    assert low.n_restart > 0                                   # max_subspace = 24 with blocks of 3
    assert low.subspace_orthogonality < 1e-9
    for val in low.eigenvalues:
        assert np.min(np.abs(spectrum - val)) < 1e-7
    np.testing.assert_allclose(high.eigenvalues, spectrum[-2:], atol=1e-7)
    # true residuals A x - theta x of the returned vectors
    for val, vec, rn in zip(low.eigenvalues, low.eigenvectors, low.residual_norms):
        r = matrix @ vec - val * vec
        assert abs(np.sqrt(r @ r) - rn) < 1e-7 and rn < 1e-5
        assert abs(vec @ vec - 1.0) < 1e-8
    with pytest.warns(la.LinAlgWarning, match="Maximum number of iterations"):
        short = lanczos(matrix, guesses, n_ep=2, which="SA", max_iter=2)
    assert not short.converged and short.n_iter == 2 and short.n_applies == 6
    with pytest.raises(TypeError):
        LanczosIterator(matrix, ["not a vector"])
    with pytest.raises(ValueError):
        LanczosIterator(matrix, guesses, ritz_vectors=guesses[:1], ritz_values=np.zeros(2),
                        ritz_overlaps=np.zeros((2, 3)))
    with pytest.raises(adcc.InputError):
        adcc.run_adc(matrix, n_singlets=2, eigensolver="arnoldi", output=None)
