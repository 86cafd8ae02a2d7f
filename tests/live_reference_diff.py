"""Differential checks against the reference's own code, run live in the build container (needs
/root/reference; executed in a subprocess by tests/test_reference_goldens_cpu.py because importing the
reference package imports libadcc, which switches select_n_* to libtensor's symmetry-unique mode):

  * adcc.workflow.validate_state_parameters vs adcc_b200.workflow.validate_state_parameters on every
    combination of n_states / n_singlets / n_triplets / n_spin_flip / kind (810 cases, restricted and
    unrestricted reference): same result tuple or same exception type;
  * adcc.guess.guesses_{singlet,triplet,any,spin_flip} vs the product's, singles and doubles blocks,
    1 ... 8 guesses, ADC(2), ADC(2)-x, CVS-ADC(2)-x: identical vectors;
  * sigma and diagonal of random vectors, the reference's AdcMatrix vs the product's, for four inputs
    (H2O / STO-3G, an unrestricted open-shell input, two restricted synthetic inputs), six choices of
    core / frozen-core / frozen-virtual spaces and all ten methods: equal to 1e-11;
  * run_adc of both packages on the same inputs (singlet / triplet / any / spin_flip, ADC(1) ... ADC(3) and
    their CVS variants, with and without frozen spaces; 116 runs): energies to 1e-8, identical iteration
    counts, same exception type where no run is possible; oscillator strengths and state dipole moments of
    the singlet runs (36 sets) to 1e-6;
  * MP2 / MP3 energies, t2, td2, mp2_diffdm (10 ground states), AdcMatrixShifted / AdcMatrixProjected, and
    the Lanczos solver (12 runs, identical iteration counts).
"""
import itertools
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools", "refshim"), REF,
                os.path.join(REF, "examples", "water")]
import cpu_shim  # noqa: E402


class _Patch:
    def setattr(self, obj, name, val, raising=True):
        setattr(obj, name, val)


cpu_shim.install(_Patch())
import adcc  # noqa: E402  (the reference package)
import adcc_b200  # noqa: E402
import adcc.guess as gref  # noqa: E402
import adcc_b200.guess as gours  # noqa: E402
import oracle  # noqa: E402
from adcc.workflow import validate_state_parameters as vref  # noqa: E402
from adcc_b200.workflow import validate_state_parameters as vours  # noqa: E402
from adc_parity_common import open_shell_dict  # noqa: E402
from import_data import import_data  # noqa: E402

assert adcc.__file__.startswith(REF)
warnings.simplefilter("ignore")
inputs = (("restricted", import_data(), oracle.hfdata.load_h2o_sto3g(), ("singlet", "triplet", "any")),
          ("unrestricted", open_shell_dict(), open_shell_dict(), ("any", "spin_flip")))

n_validation = 0
for label, dref, dours, _ in inputs:
    r1, r2 = adcc.ReferenceState(dref), adcc_b200.ReferenceState(dours)
    vals = [None, 0, 2]
    for n_states, n_singlets, n_triplets, n_spin_flip in itertools.product(vals, vals, vals, vals):
        for kind in ("any", "singlet", "triplet", "spin_flip", "bogus"):
            def run(fn, ref):
                try:
                    return ("ok",) + tuple(fn(ref, n_states=n_states, n_singlets=n_singlets, n_triplets=n_triplets,
                                              n_spin_flip=n_spin_flip, kind=kind))
                except Exception as e:
                    return ("exc", type(e).__name__)
            a, b = run(vref, r1), run(vours, r2)
            assert a == b, (label, n_states, n_singlets, n_triplets, n_spin_flip, kind, a, b)
            n_validation += 1

n_guesses = 0
for label, dref, dours, kinds in inputs:
    for method, core in (("adc2", None), ("adc2x", None), ("cvs-adc2x", 1)):
        if label == "unrestricted" and core:
            continue
        mr = adcc.AdcMatrix(method, adcc.ReferenceState(dref, core_orbitals=core))
        mo = adcc_b200.AdcMatrix(method, adcc_b200.ReferenceState(dours, core_orbitals=core))
        for kind in kinds:
            for block in ("ph", "pphh"):
                for n in (1, 2, 3, 5, 8):
                    def make(mod, m):
                        try:
                            return getattr(mod, "guesses_" + kind)(m, n, block=block)
                        except Exception as e:
                            return type(e).__name__
                    vr, vo = make(gref, mr), make(gours, mo)
                    if isinstance(vr, str) or isinstance(vo, str):
                        assert vr == vo, (label, method, kind, block, n, vr, vo)
                    else:
                        assert len(vr) == len(vo), (label, method, kind, block, n, len(vr), len(vo))
                        for a, b in zip(vr, vo):
                            for blk in a.blocks:
                                assert np.max(np.abs(a[blk].to_ndarray() - b[blk].to_ndarray())) < 1e-12, \
                                    (label, method, kind, block, n, blk)
                    n_guesses += 1

# ---- sigma / diagonal of random vectors: reference AdcMatrix vs product AdcMatrix ------------------
from adcc.guess import guess_zero  # noqa: E402
from adcc_b200.synthetic import restricted_scf_dict  # noqa: E402


def with_spin_orbital_eri(data):
    """adcc_b200.synthetic dicts keep (pq|rs) spatial; the reference's DataHfProvider wants eri_ffff."""
    d = dict(data)
    g = d.pop("eri_ffff_spatial")
    n = g.shape[0]
    E = np.zeros((2 * n,) * 4)
    for a in (slice(0, n), slice(n, 2 * n)):
        for b in (slice(0, n), slice(n, 2 * n)):
            E[a, a, b, b] = g
    d["eri_ffff"] = E
    return d


small, mid = restricted_scf_dict(5, 2, seed=21), restricted_scf_dict(7, 3, seed=22)
sigma_inputs = [("h2o", import_data(), oracle.hfdata.load_h2o_sto3g()), ("open-shell", open_shell_dict(), open_shell_dict()),
                ("synthetic 5/2", with_spin_orbital_eri(small), small), ("synthetic 7/3", with_spin_orbital_eri(mid), mid)]
spaces = [dict(), dict(core_orbitals=1), dict(frozen_core=1), dict(frozen_virtual=1),
          dict(core_orbitals=1, frozen_virtual=1), dict(frozen_core=1, frozen_virtual=2)]
n_sigma = n_packed = 0
for label, dref, dours in sigma_inputs:
    for kw in spaces:
        if label == "open-shell" and kw:
            continue                                   # subspace lists of unequal alpha / beta counts: not probed here
        if label == "synthetic 5/2" and kw.get("frozen_virtual") == 2 and kw.get("frozen_core"):
            continue                                   # would leave one active occupied and one virtual pair
        cvs = "core_orbitals" in kw
        for method in (("cvs-adc0", "cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3") if cvs else
                       ("adc0", "adc1", "adc2", "adc2x", "adc3")):
            try:
                mo = adcc_b200.AdcMatrix(method, adcc_b200.ReferenceState(dours, **kw))
            except ValueError as e:                   # e.g. no virtual orbital left: the reference must object too
                try:
                    adcc.AdcMatrix(method, adcc.ReferenceState(dref, **kw))
                except ValueError:
                    continue
                raise AssertionError((label, kw, method, "only the product raised", str(e)))
            mr = adcc.AdcMatrix(method, adcc.ReferenceState(dref, **kw))
            rng = np.random.default_rng(n_sigma + 5)
            vr = guess_zero(mr)
            blocks = {}
            for blk in vr.blocks:
                arr = rng.standard_normal(vr[blk].shape)
                if blk == "pphh":
                    arr = arr - arr.transpose(0, 1, 3, 2)
                    if not cvs:
                        arr = arr - arr.transpose(1, 0, 2, 3)
                vr[blk].set_from_ndarray(arr)
                blocks[blk] = adcc_b200.Tensor.from_ndarray(mo.mospaces, mo.axis_spaces[blk], arr)
            sr = adcc.evaluate(mr @ vr)
            so = mo.matvec(adcc_b200.AmplitudeVector(**blocks))
            for blk in vr.blocks:
                a, b = sr[blk].to_ndarray(), so[blk].to_ndarray()
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(a)), 1e-300), (label, kw, method, blk)
                a, b = mr.diagonal()[blk].to_ndarray(), mo.diagonal()[blk].to_ndarray()
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(a)), 1e-300), (label, kw, method, "diag", blk)
            n_sigma += 1
            # the hand-scheduled engines on the packed stores, same vector
            if method in ("adc2", "adc2x", "adc3", "cvs-adc2", "cvs-adc2x"):
                from adcc_b200.packed import PackedDoubles
                from adcc_b200.packed_cvs import CvsPackedDoubles
                mp_ = adcc_b200.AdcMatrix(method, adcc_b200.ReferenceState(dours, **kw), doubles_storage="packed")
                assert mp_.doubles_storage == "packed"
                pb = dict(blocks)
                pb["pphh"] = (CvsPackedDoubles if cvs else PackedDoubles).from_dense(blocks["pphh"])
                sp = mp_.matvec(adcc_b200.AmplitudeVector(**pb))
                for blk in vr.blocks:
                    a, b = sr[blk].to_ndarray(), sp[blk].to_ndarray()
                    assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(a)), 1e-300), (label, kw, method, "packed", blk)
                n_packed += 1

# ---- converged states: reference run_adc vs product run_adc -----------------------------------------
n_runs = n_runs_packed = n_props = 0
for label, dref, dours in sigma_inputs:
    restricted = label != "open-shell"
    for kw in (dict(), dict(core_orbitals=1), dict(frozen_core=1, frozen_virtual=1)):
        if not restricted and kw:
            continue
        cvs = "core_orbitals" in kw
        for method in (("cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3") if cvs else ("adc1", "adc2", "adc2x", "adc3")):
            for kind in (("singlet", "triplet", "any") if restricted else ("any", "spin_flip")):
                key = {"singlet": "n_singlets", "triplet": "n_triplets", "any": "n_states", "spin_flip": "n_spin_flip"}[kind]
                res = []
                for mod, d in ((adcc, dref), (adcc_b200, dours)):
                    try:
                        st = mod.run_adc(d, method=method, conv_tol=1e-8, output=None, **{key: 2}, **kw)
                        res.append((bool(st.converged), int(st.n_iter), np.array(st.excitation_energy), st))
                    except Exception as e:           # e.g. fewer guesses than requested: both must agree
                        res.append(type(e).__name__)
                a, b = res
                if isinstance(a, str) or isinstance(b, str):
                    assert a == b, (label, kw, method, kind, a, b)
                else:
                    assert a[0] and b[0], (label, kw, method, kind, "not converged")
                    assert np.max(np.abs(a[2] - b[2])) < 1e-8, (label, kw, method, kind, a[2], b[2])
                    assert a[1] == b[1], (label, kw, method, kind, "iterations", a[1], b[1])
                    # the plain table (energies, block norms, header with kind / spin change) of every run
                    has_dip = restricted
                    ta = a[3].describe(oscillator_strengths=has_dip).replace("-0.0000", " 0.0000")
                    tb = b[3].describe(oscillator_strengths=has_dip).replace("-0.0000", " 0.0000")
                    assert ta == tb, (label, kw, method, kind, "describe()\n" + ta + "\n" + tb)
                    if kind == "singlet":
                        # the reference's property code vs adcc_b200/properties.py on the same states
                        fa, fb = np.array(a[3].oscillator_strength), np.array(b[3].oscillator_strength)
                        assert np.max(np.abs(fa - fb)) < 1e-6 * max(1.0, np.max(np.abs(fa))), (label, kw, method, "f")
                        da, db = np.array(a[3].state_dipole_moment), np.array(b[3].state_dipole_moment)
                        assert np.max(np.abs(da - db)) < 1e-6, (label, kw, method, "state dipole moment")
                        # density matrices block by block (state densities are quadratic in the vector: no sign;
                        # transition densities up to the sign of the eigenvector)
                        full = {"o": "o1", "v": "v1", "c": "o2"}
                        for n in range(2):
                            ddr, ddo = a[3].state_diffdm[n], b[3].state_diffdm[n]
                            for blk, rho in ddo.items():
                                name = full[blk[0]] + full[blk[1]]
                                want = ddr[name].to_ndarray() if name in ddr.blocks_nonzero else 0.0 * rho.to_ndarray()
                                assert np.max(np.abs(rho.to_ndarray() - want)) < 1e-6, (label, kw, method, "diffdm", blk)
                            tdr, tdo = a[3].transition_dm[n], b[3].transition_dm[n]
                            big = max(tdo, key=lambda k: float(np.max(np.abs(tdo[k].to_ndarray()))))
                            ref_big = tdr[full[big[0]] + full[big[1]]].to_ndarray()
                            k = np.unravel_index(np.argmax(np.abs(ref_big)), ref_big.shape)
                            sign = 1.0 if tdo[big].to_ndarray()[k] * ref_big[k] >= 0 else -1.0
                            for blk, rho in tdo.items():
                                name = full[blk[0]] + full[blk[1]]
                                want = tdr[name].to_ndarray() if name in tdr.blocks_nonzero else 0.0 * rho.to_ndarray()
                                assert np.max(np.abs(sign * rho.to_ndarray() - want)) < 1e-6, (label, kw, method, "tdm", blk)
                        if not cvs:                  # transitions between the two excited states
                            sa = np.array(adcc.State2States(a[3], initial=0).oscillator_strength)
                            sb = np.array(adcc_b200.State2States(b[3], initial=0).oscillator_strength)
                            assert np.max(np.abs(sa - sb)) < 1e-6 * max(1.0, np.max(np.abs(sa))), (label, kw, method, "s2s")
                        # the table users read (adcc/ExcitedStates.py:50-166)
                        ta = a[3].describe(state_dipole_moments=True, transition_dipole_moments=False)
                        tb = b[3].describe(state_dipole_moments=True, transition_dipole_moments=False)
                        ta, tb = ta.replace("-0.0000", " 0.0000"), tb.replace("-0.0000", " 0.0000")   # sign of a 1e-17
                        assert ta == tb, (label, kw, method, "describe()\n" + ta + "\n" + tb)
                        n_props += 1
                    if method in ("adc2", "adc2x", "adc3", "cvs-adc2", "cvs-adc2x"):
                        mp_ = adcc_b200.AdcMatrix(method, adcc_b200.ReferenceState(dours, **kw), doubles_storage="packed")
                        sp = adcc_b200.run_adc(mp_, conv_tol=1e-8, output=None, **{key: 2})
                        assert sp.converged and np.max(np.abs(a[2] - sp.excitation_energy)) < 1e-8, \
                            (label, kw, method, kind, "packed")
                        assert sp.n_iter == a[1], (label, kw, method, kind, "packed iterations", sp.n_iter, a[1])
                        n_runs_packed += 1
                n_runs += 1

# ---- MP ground state quantities, shifted / projected matrices, Lanczos ------------------------------
def close(a, b, what, tol=1e-11):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape and np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(a)), 1e-300), what


n_mp = n_wrapped = n_lanczos = n_intermediates = n_dense = 0
for label, dref, dours in sigma_inputs:
    restricted = label != "open-shell"
    for kw in ((dict(), dict(core_orbitals=1), dict(frozen_core=1, frozen_virtual=1)) if restricted else (dict(),)):
        rr, ro = adcc.ReferenceState(dref, **kw), adcc_b200.ReferenceState(dours, **kw)
        gr, go = adcc.LazyMp(rr), adcc_b200.LazyMp(ro)
        close(gr.energy_correction(2), go.energy_correction(2), (label, kw, "MP2"))
        if "core_orbitals" not in kw:
            close(gr.energy_correction(3), go.energy_correction(3), (label, kw, "MP3"))
            close(gr.td2("o1o1v1v1").to_ndarray(), go.td2("o1o1v1v1").to_ndarray(), (label, kw, "td2"))
        close(gr.t2("o1o1v1v1").to_ndarray(), go.t2("o1o1v1v1").to_ndarray(), (label, kw, "t2"))
        dr, do = gr.mp2_diffdm, go.mp2_diffdm
        for blk, name in (("o1o1", "oo"), ("o1v1", "ov"), ("v1v1", "vv")):
            close(dr[blk].to_ndarray(), do[name].to_ndarray(), (label, kw, "mp2_diffdm", blk))
        n_mp += 1
    # ADC(2) / ADC(3) intermediates (adcc/adc_pp/matrix.py:933-1094) of both packages
    ir = adcc.AdcMatrix("adc3", adcc.ReferenceState(dref)).intermediates
    io = adcc_b200.AdcMatrix("adc3", adcc_b200.ReferenceState(dours), doubles_storage="dense").intermediates
    for name in ("adc2_i1", "adc2_i2", "adc3_m11", "adc3_pia", "adc3_pib", "t2sq"):
        close(getattr(ir, name).to_ndarray(), getattr(io, name).to_ndarray(), (label, "intermediate", name))
        n_intermediates += 1
    # dense export in the antisymmetrised basis and block views (adcc/AdcMatrix.py:526-646)
    if label in ("h2o", "synthetic 5/2"):
        for method, kwx in (("adc1", {}), ("adc2", {}), ("cvs-adc2x", dict(core_orbitals=1))):
            er = adcc.AdcMatrix(method, adcc.ReferenceState(dref, **kwx))
            eo = adcc_b200.AdcMatrix(method, adcc_b200.ReferenceState(dours, **kwx))
            close(er.to_ndarray(), eo.to_ndarray(), (label, method, "to_ndarray"))
            assert [tuple(b) for b in er.dense_basis()] == [tuple(b) for b in eo.dense_basis()], (label, method, "dense_basis")
            close(er.block_view("ph_ph").to_ndarray(), eo.block_view("ph_ph").to_ndarray(), (label, method, "block_view"))
            n_dense += 1
    # shifted and projected matrices on a random vector (adcc/AdcMatrix.py:649-780)
    from adcc.AdcMatrix import AdcMatrixProjected as RefProjected, AdcMatrixShifted as RefShifted
    mr, mo = adcc.AdcMatrix("adc2x", adcc.ReferenceState(dref)), adcc_b200.AdcMatrix("adc2x", adcc_b200.ReferenceState(dours))
    rng = np.random.default_rng(99)
    vr = guess_zero(mr)
    blocks = {}
    for blk in vr.blocks:
        arr = rng.standard_normal(vr[blk].shape)
        if blk == "pphh":
            arr = arr - arr.transpose(0, 1, 3, 2)
            arr = arr - arr.transpose(1, 0, 2, 3)
        vr[blk].set_from_ndarray(arr)
        blocks[blk] = adcc_b200.Tensor.from_ndarray(mo.mospaces, mo.axis_spaces[blk], arr)
    vo = adcc_b200.AmplitudeVector(**blocks)
    pairs = [(RefShifted(mr, -0.4), adcc_b200.AdcMatrixShifted(mo, -0.4))]
    if restricted:
        pairs.append((RefProjected(mr, ["cv", "ccvv", "ocvv"], core_orbitals=1),
                      adcc_b200.AdcMatrixProjected(mo, ["cv", "ccvv", "ocvv"], core_orbitals=1)))
    for wr, wo in pairs:
        sr, so = adcc.evaluate(wr @ vr), wo.matvec(vo)
        for blk in vr.blocks:
            close(sr[blk].to_ndarray(), so[blk].to_ndarray(), (label, type(wo).__name__, blk))
            close(wr.diagonal()[blk].to_ndarray(), wo.diagonal()[blk].to_ndarray(), (label, type(wo).__name__, "diag", blk))
        n_wrapped += 1
    # the second eigensolver
    if restricted:
        for method in ("adc2", "adc2x"):
            for key in ("n_singlets", "n_triplets"):
                res = []
                for mod, d in ((adcc, dref), (adcc_b200, dours)):
                    st = mod.run_adc(d, method=method, conv_tol=1e-7, output=None, eigensolver="lanczos", **{key: 2})
                    res.append((bool(st.converged), int(st.n_iter), np.array(st.excitation_energy)))
                a, b = res
                assert a[0] == b[0] and a[1] == b[1], (label, method, key, "lanczos", a[:2], b[:2])
                if a[0]:
                    assert np.max(np.abs(a[2] - b[2])) < 1e-7, (label, method, key, "lanczos energies")
                n_lanczos += 1
print(f"ok {n_validation} validation cases, {n_guesses} guess sets, {n_sigma} sigma / diagonal comparisons, {n_packed} on the packed engines, {n_runs} Davidson runs, {n_runs_packed} also on the packed engines, {n_props} property sets, {n_mp} MP ground states, {n_wrapped} shifted / projected matrices, {n_lanczos} Lanczos runs, {n_intermediates} intermediates, {n_dense} dense exports")
