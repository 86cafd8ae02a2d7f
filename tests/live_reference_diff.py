"""Differential checks against the reference's own code, run live in the build container (needs
/root/reference; executed in a subprocess by tests/test_reference_goldens_cpu.py because importing the
reference package imports libadcc, which switches select_n_* to libtensor's symmetry-unique mode):

  * adcc.workflow.validate_state_parameters vs adcc_b200.workflow.validate_state_parameters on every
    combination of n_states / n_singlets / n_triplets / n_spin_flip / kind (810 cases, restricted and
    unrestricted reference): same result tuple or same exception type;
  * adcc.guess.guesses_{singlet,triplet,any,spin_flip} vs the product's, singles and doubles blocks,
    1 ... 8 guesses, ADC(2), ADC(2)-x, CVS-ADC(2)-x: identical vectors.
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
print(f"ok {n_validation} validation cases, {n_guesses} guess sets")
