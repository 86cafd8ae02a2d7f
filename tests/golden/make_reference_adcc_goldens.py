#!/usr/bin/env python3
"""Golden sigma vectors / diagonals / block results / energies produced by the REFERENCE's own
Python layer: the unmodified ``adcc`` package of /root/reference (workflow, AdcMatrix,
adc_pp/matrix.py, LazyMp, solver/davidson.py, guess/*) imported in the build container and run on
top of this repository's ``libadcc`` module surface (libadcc/__init__.py), with

  * the reference's own fixture examples/water/data.py (H2O / STO-3G RHF) as input,
  * the tensor arithmetic done by torch CPU kernels (tests/cpu_shim.py) - no CUDA here -, and
  * tools/refshim/ standing in for the absent third-party packages opt_einsum / h5py.

What this pins: the EQUATIONS and the solver control flow are the reference's, executed, not
transcribed.  The dump mirrors the reference's own generator
adcc/tests/generators/dump_adcc.py:192-263 (random trial vector, matvec, diagonal, per-block
results) and is compared by tests/test_reference_goldens_*.py with the oracle, the product host
path and the CUDA path at the tolerances of adcc/tests/AdcMatrix_test.py:213-285 or tighter.

    python tests/golden/make_reference_adcc_goldens.py      # writes reference_adcc_h2o_sto3g.npz and the
                                                            # open_shell / workflows / baseline_shapes files
    python tests/golden/make_reference_adcc_goldens.py baseline_shapes      # one of the latter only
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools", "refshim"), REF,
                os.path.join(REF, "examples", "water")]

OPEN_SHELL_METHODS = ["adc1", "adc2", "adc2x", "adc3"]
LANCZOS_METHODS = ["adc2", "adc2x", "cvs-adc2"]      # ADC(0) / ADC(1) of this fixture exhaust their Krylov space
METHODS = ["adc0", "adc1", "adc2", "adc2x", "adc3", "cvs-adc0", "cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3"]


def main():
    import cpu_shim

    class Patch:
        def setattr(self, obj, name, val, raising=True):
            setattr(obj, name, val)
    cpu_shim.install(Patch())
    import adcc                                  # the reference package, unmodified
    from adcc.guess import guess_zero
    from import_data import import_data          # examples/water/import_data.py
    assert adcc.__file__.startswith(REF), adcc.__file__

    data = import_data()
    only = sys.argv[1:]                           # e.g. "baseline_shapes": regenerate one file only
    if only:
        parts = {"open_shell": lambda: open_shell(adcc, guess_zero), "workflows": lambda: workflows(adcc, data),
                 "baseline_shapes": lambda: baseline_shapes(adcc, guess_zero),
                 "synthetic_workload": lambda: synthetic_workload(adcc, guess_zero)}
        for name in only:
            parts[name]()
        return
    out = {"methods": np.array(METHODS)}
    for method in METHODS:
        cvs = method.startswith("cvs")
        ref = adcc.ReferenceState(data, core_orbitals=1 if cvs else None)
        matrix = adcc.AdcMatrix(method, ref)
        key = method.replace("-", "_")
        rng = np.random.default_rng(sum(map(ord, key)))
        vec = guess_zero(matrix)
        for blk in vec.blocks:
            arr = rng.standard_normal(vec[blk].shape)
            if blk == "pphh":                     # antisymmetric like every doubles vector
                arr = arr - arr.transpose(0, 1, 3, 2)
                if not cvs:
                    arr = arr - arr.transpose(1, 0, 2, 3)
            vec[blk].set_from_ndarray(arr)
            out[f"{key}/vector/{blk}"] = arr
        sigma = adcc.evaluate(matrix @ vec)       # AdcMatrix.matvec, adcc/AdcMatrix.py:390-396
        diag = matrix.diagonal()
        for blk in vec.blocks:
            out[f"{key}/matvec/{blk}"] = sigma[blk].to_ndarray()
            out[f"{key}/diagonal/{blk}"] = diag[blk].to_ndarray()
        for block in matrix.blocks:               # AdcMatrix.block_apply, adcc/AdcMatrix.py:375-388
            outb, inb = block.split("_")
            if inb not in vec.blocks or outb not in vec.blocks:
                continue
            res = matrix.block_apply(block, vec[inb])
            out[f"{key}/block/{block}"] = adcc.evaluate(res).to_ndarray()
        if not cvs:
            mp = matrix.ground_state
            out[f"{key}/mp2"] = np.array(mp.energy_correction(2))
            if method == "adc3":
                out[f"{key}/mp3"] = np.array(mp.energy_correction(3))
        for kind, kw in (("singlet", "n_singlets"), ("triplet", "n_triplets")):
            n = 2 if cvs else 3
            state = adcc.run_adc(matrix, **{kw: n}, conv_tol=1e-9)
            assert state.converged
            out[f"{key}/{kind}/energies"] = np.array(state.excitation_energy)
            out[f"{key}/{kind}/n_iter"] = np.array(state.n_iter)
            if kind == "singlet" and method not in ("adc0", "cvs-adc0"):
                # the reference's own property code: adc_pp/transition_dm.py, OperatorIntegrals.py,
                # ElectronicTransition._oscillator_strength (2/3 |mu|^2 omega)
                out[f"{key}/singlet/oscillator_strength"] = np.array(state.oscillator_strength)
                out[f"{key}/singlet/transition_dipole_moment"] = np.array(state.transition_dipole_moment)
            if kind == "singlet":
                # the reference's state densities: adc_pp/state_diffdm.py, ElectronicStates.state_dipole_moment,
                # GroundState.dipole_moment (the dense matrix is in the order of the full orbital axis:
                # OneParticleDensity.to_ndarray, adcc/NParticleOperator.py)
                out[f"{key}/singlet/state_dipole_moment"] = np.array(state.state_dipole_moment)
                level = state.property_method.level.to_int()
                gs = state.reference_state.dipole_moment if level == 0 else state.ground_state.dipole_moment(level)
                out[f"{key}/ground_state_dipole_moment"] = np.array(gs)
                for n, ddm in enumerate(state.state_diffdm):
                    for blk in ddm.blocks_nonzero:
                        out[f"{key}/singlet/state_diffdm/{n}/{blk}"] = ddm[blk].to_ndarray()
                if not cvs:
                    # transitions between excited states: adcc/State2States.py, adc_pp/state2state_transition_dm.py
                    s2s = adcc.State2States(state, initial=0)
                    out[f"{key}/singlet/s2s/excitation_energy"] = np.array(s2s.excitation_energy)
                    out[f"{key}/singlet/s2s/transition_dipole_moment"] = np.array(s2s.transition_dipole_moment)
                    out[f"{key}/singlet/s2s/oscillator_strength"] = np.array(s2s.oscillator_strength)
                    for n, tdm in enumerate(s2s.transition_dm):
                        for blk in tdm.blocks_nonzero:
                            out[f"{key}/singlet/s2s/transition_dm/{n}/{blk}"] = tdm[blk].to_ndarray()
        if method in LANCZOS_METHODS:
            # the reference's second eigensolver (adcc/solver/lanczos.py, LanczosIterator.py, orthogonaliser.py)
            # through run_adc(..., eigensolver="lanczos")
            for kind, kw in (("singlet", "n_singlets"), ("triplet", "n_triplets")):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    state = adcc.run_adc(matrix, **{kw: 2 if cvs else 3}, conv_tol=1e-8, eigensolver="lanczos",
                                         output=None)
                assert state.converged
                out[f"{key}/lanczos/{kind}/energies"] = np.array(state.excitation_energy)
                out[f"{key}/lanczos/{kind}/n_iter"] = np.array(state.n_iter)
        print(method, out[f"{key}/singlet/energies"], out[f"{key}/singlet/n_iter"], flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_adcc_h2o_sto3g.npz"), **out)
    open_shell(adcc, guess_zero)
    workflows(adcc, data)
    baseline_shapes(adcc, guess_zero)
    synthetic_workload(adcc, guess_zero)


def workflows(adcc, data):
    """Callers around the matvec that BASELINE's configs use: the projected ADC(2)-x matrix of
    examples/2_core_excitations/water_core_by_projection.py (AdcMatrixProjected, CVS guesses
    transferred to the full space) and the reference's frozen-core / frozen-virtual subspace cases
    (adcc/tests/testcases.py: fc, fv, fc-fv, fv-cvs) on the H2O / STO-3G input."""
    import adcc.projection
    from adcc.AdcMatrix import AdcMatrixProjected
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        matrix = adcc.AdcMatrix("adc2x", adcc.ReferenceState(data))
        pmatrix = AdcMatrixProjected(matrix, ["cv", "ccvv", "ocvv"], core_orbitals=1)
        sp = adcc.run_adc(pmatrix, n_singlets=2, conv_tol=1e-8, n_guesses=2, n_guesses_doubles=0, output=None)
        sc = adcc.cvs_adc2x(data, n_singlets=2, conv_tol=1e-8, core_orbitals=1, output=None)
        guesses = adcc.projection.transfer_cvs_to_full(sc, pmatrix)
        sp2 = adcc.run_adc(pmatrix, n_singlets=2, conv_tol=1e-8, guesses=guesses, output=None)
    assert sp.converged and sp2.converged and sc.converged
    out["projected/energies"], out["projected/n_iter"] = np.array(sp.excitation_energy), np.array(sp.n_iter)
    out["projected_from_cvs/energies"] = np.array(sp2.excitation_energy)
    out["projected_from_cvs/n_iter"] = np.array(sp2.n_iter)
    out["projected/diagonal/ph"] = pmatrix.diagonal().ph.to_ndarray()
    out["projected/diagonal/pphh"] = pmatrix.diagonal().pphh.to_ndarray()
    for name, method, kw in (("fc", "adc2", dict(frozen_core=1)), ("fv", "adc2", dict(frozen_virtual=1)),
                             ("fc_fv", "adc2", dict(frozen_core=1, frozen_virtual=1)),
                             ("fv_cvs", "cvs-adc2", dict(core_orbitals=1, frozen_virtual=1)),
                             ("fc_adc3", "adc3", dict(frozen_core=1))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = adcc.run_adc(data, method=method, n_singlets=2, conv_tol=1e-8, output=None, **kw)
        assert st.converged
        out[f"{name}/energies"], out[f"{name}/n_iter"] = np.array(st.excitation_energy), np.array(st.n_iter)
        if method != "adc3":
            out[f"{name}/oscillator_strength"] = np.array(st.oscillator_strength)
        print("workflow", name, out[f"{name}/energies"], out[f"{name}/n_iter"], flush=True)
    print("workflow projected", out["projected/energies"], out["projected/n_iter"], out["projected_from_cvs/n_iter"])
    np.savez_compressed(os.path.join(HERE, "reference_adcc_workflows.npz"), **out)


def open_shell(adcc, guess_zero):
    """Second input: the UNRESTRICTED open-shell SCF dict of tests/adc_parity_common.py::open_shell_dict
    (3 alpha / 2 beta electrons, 6 spatial orbitals, <pq||rs> given directly, no spin-block
    structure in the integrals) - the reference's unrestricted code paths (no spin-block maps,
    ragged alpha / beta halves on every axis, kind "any" and "spin_flip" guesses)."""
    from adc_parity_common import open_shell_dict
    data = open_shell_dict()
    out = {"methods": np.array(OPEN_SHELL_METHODS)}
    for method in OPEN_SHELL_METHODS:
        ref = adcc.ReferenceState(data)
        matrix = adcc.AdcMatrix(method, ref)
        rng = np.random.default_rng(sum(map(ord, method)) + 7)
        vec = guess_zero(matrix)
        for blk in vec.blocks:
            arr = rng.standard_normal(vec[blk].shape)
            if blk == "pphh":
                arr = arr - arr.transpose(0, 1, 3, 2)
                arr = arr - arr.transpose(1, 0, 2, 3)
            vec[blk].set_from_ndarray(arr)
            out[f"{method}/vector/{blk}"] = arr
        sigma = adcc.evaluate(matrix @ vec)
        diag = matrix.diagonal()
        for blk in vec.blocks:
            out[f"{method}/matvec/{blk}"] = sigma[blk].to_ndarray()
            out[f"{method}/diagonal/{blk}"] = diag[blk].to_ndarray()
        out[f"{method}/mp2"] = np.array(matrix.ground_state.energy_correction(2))
        for kind, kw, n in (("any", "n_states", 3), ("spin_flip", "n_spin_flip", 2)):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                state = adcc.run_adc(matrix, **{kw: n}, conv_tol=1e-8, output=None)
            assert state.converged
            out[f"{method}/{kind}/energies"] = np.array(state.excitation_energy)
            out[f"{method}/{kind}/n_iter"] = np.array(state.n_iter)
        print("open shell", method, out[f"{method}/any/energies"], out[f"{method}/any/n_iter"],
              out[f"{method}/spin_flip/n_iter"], flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_adcc_open_shell.npz"), **out)




BASELINE_SHAPE_CASES = {     # key: (named_case, method, run_adc arguments as in BASELINE.json's configs)
    "c1": ("h2o_ccpvdz", "adc2", dict(n_singlets=3, conv_tol=1e-6)),
    "c1_adc3": ("h2o_ccpvdz", "adc3", dict(n_singlets=3, conv_tol=1e-6)),
    "c2": ("h2o_ccpvtz", "adc2", dict(n_singlets=10)),
    "c4": ("h2o_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=7, conv_tol=1e-8)),
    "c4_ne": ("ne_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=4, conv_tol=1e-8)),
    "c3": ("methyloxirane_ccpvdz", "adc3", dict(n_singlets=8)),       # hours of CPU: only with REF_GOLDEN_C3=1
}


def spin_orbital_eri(data):
    """The SCF dicts of adcc_b200.synthetic carry the spatial (pq|rs) under an extension key; the
    reference's DataHfProvider wants the full spin-orbital chemists' array (DataHfProvider.py:122-132)."""
    d = dict(data)
    g = d.pop("eri_ffff_spatial")
    n = g.shape[0]
    E = np.zeros((2 * n,) * 4)
    for a in (slice(0, n), slice(n, 2 * n)):
        for b in (slice(0, n), slice(n, 2 * n)):
            E[a, a, b, b] = g
    d["eri_ffff"] = E
    return d


def baseline_shapes(adcc, guess_zero):
    """The reference's own run_adc on the restricted synthetic SCF inputs of the exact BASELINE
    shapes (adcc_b200.synthetic.named_case: configs[0..3]): energies, oscillator strengths,
    iteration counts, and sigma of a seeded trial vector (singles complete, 4096 sampled doubles)."""
    from adcc_b200.synthetic import named_case
    path = os.path.join(HERE, "reference_adcc_baseline_shapes.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for key, (name, method, kw) in BASELINE_SHAPE_CASES.items():
        if key == "c3" and os.environ.get("REF_GOLDEN_C3") != "1":
            continue
        data, core = named_case(name)
        ref = adcc.ReferenceState(spin_orbital_eri(data), core_orbitals=core or None)
        matrix = adcc.AdcMatrix(method, ref)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            state = adcc.run_adc(matrix, output=None, **kw)
        assert state.converged
        out[f"{key}/energies"], out[f"{key}/n_iter"] = np.array(state.excitation_energy), np.array(state.n_iter)
        out[f"{key}/oscillator_strength"] = np.array(state.oscillator_strength)
        rng = np.random.default_rng(sum(map(ord, key)))
        vec = guess_zero(matrix)
        cvs = method.startswith("cvs")
        for blk in vec.blocks:
            arr = rng.standard_normal(vec[blk].shape)
            if blk == "pphh":
                arr = arr - arr.transpose(0, 1, 3, 2)
                if not cvs:
                    arr = arr - arr.transpose(1, 0, 2, 3)
            vec[blk].set_from_ndarray(arr)
        sigma = adcc.evaluate(matrix @ vec)
        out[f"{key}/vector_seed"] = np.array(sum(map(ord, key)))
        out[f"{key}/matvec/ph"] = sigma.ph.to_ndarray()
        s2 = sigma.pphh.to_ndarray().reshape(-1)
        pick = np.random.default_rng(1).choice(s2.size, size=min(4096, s2.size), replace=False)
        out[f"{key}/matvec/pphh_index"], out[f"{key}/matvec/pphh_value"] = pick, s2[pick]
        out[f"{key}/matvec/pphh_max"] = np.array(np.max(np.abs(s2)))
        print("baseline shape", key, method, out[f"{key}/energies"][:3], out[f"{key}/n_iter"], flush=True)
        del matrix, ref, state, sigma, vec
    np.savez_compressed(path, **out)


SYNTHETIC_CASES = [(6, 10, 5), (8, 20, 11)]        # (nocc, nvirt, seed) of the counter-based workload


def synthetic_workload(adcc, guess_zero):
    """BASELINE configs[4]'s input definition at small sizes: the counter-based <pq||rs> of
    adcc_b200/synthetic.py written out as a plain SCF dict (tests/adc_parity_common.py::
    synthetic_workload_dict) and handed to the reference's own ReferenceState / AdcMatrix / run_adc:
    sigma of a seeded trial vector, the diagonal and three states of ADC(2), ADC(2)-x and ADC(3)."""
    from adc_parity_common import synthetic_workload_dict
    out = {"cases": np.array(SYNTHETIC_CASES)}
    for no, nv, seed in SYNTHETIC_CASES:
        data = synthetic_workload_dict(no, nv, seed, 0.02)
        for method in ("adc2", "adc2x", "adc3"):
            key = f"{no}_{nv}/{method}"
            ref = adcc.ReferenceState(data)
            matrix = adcc.AdcMatrix(method, ref)
            rng = np.random.default_rng(no * 1000 + nv + sum(map(ord, method)))
            vec = guess_zero(matrix)
            for blk in vec.blocks:
                arr = rng.standard_normal(vec[blk].shape)
                if blk == "pphh":
                    arr = arr - arr.transpose(0, 1, 3, 2)
                    arr = arr - arr.transpose(1, 0, 2, 3)
                vec[blk].set_from_ndarray(arr)
                out[f"{key}/vector/{blk}"] = arr
            sigma = adcc.evaluate(matrix @ vec)
            diag = matrix.diagonal()
            for blk in vec.blocks:
                out[f"{key}/matvec/{blk}"] = sigma[blk].to_ndarray()
                out[f"{key}/diagonal/{blk}"] = diag[blk].to_ndarray()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                state = adcc.run_adc(matrix, n_states=3, conv_tol=1e-8, output=None)
            assert state.converged
            out[f"{key}/energies"], out[f"{key}/n_iter"] = np.array(state.excitation_energy), np.array(state.n_iter)
            print("synthetic workload", key, out[f"{key}/energies"], out[f"{key}/n_iter"], flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_adcc_synthetic_workload.npz"), **out)


if __name__ == "__main__":
    main()
