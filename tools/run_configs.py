"""Time-to-converged states for the molecular-shape configs of BASELINE.json (synthetic restricted
SCF data of the same shapes), through run_adc.  ADCCB_STORAGE / ADCCB_CVS_STORAGE
(auto | dense | packed) choose the doubles store of the generic / CVS cases."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adcc_b200 as adcc
from adcc_b200 import _cabi
from adcc_b200.synthetic import named_case

STORAGE = os.environ.get("ADCCB_STORAGE", "auto")      # auto | dense | packed (ADC(2), ADC(2)-x, ADC(3))
CASES = [("h2o_ccpvdz", "adc2", dict(n_singlets=3, conv_tol=1e-6)),
         ("h2o_ccpvtz", "adc2", dict(n_singlets=10)),
         ("methyloxirane_ccpvdz", "adc3", dict(n_singlets=8)),
         ("ne_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=4, conv_tol=1e-8)),
         ("h2o_ccpvtz_cvs", "cvs-adc2x", dict(n_singlets=7, conv_tol=1e-8))]
# beyond BASELINE: a CVS shape where the doubles store matters (c=2, o=22, v=196; only when named)
EXTRA = {"cvs_large": ("cvs-adc2x", dict(n_singlets=5, conv_tol=1e-8), (110, 12, 1))}
only = sys.argv[1:]
CASES += [(k, v[0], v[1]) for k, v in EXTRA.items() if k in only]
res = []
for name, method, kw in CASES:
    if only and name not in only:
        continue
    if name in EXTRA:
        from adcc_b200.synthetic import restricted_scf_dict
        n_orb, n_occ, core = EXTRA[name][2]
        data = restricted_scf_dict(n_orb, n_occ, seed=77)
    else:
        data, core = named_case(name)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = adcc.ReferenceState(data, core_orbitals=core or None)
    storage = os.environ.get("ADCCB_CVS_STORAGE", "auto") if method.startswith("cvs") else STORAGE
    matrix = adcc.AdcMatrix(method, ref, doubles_storage=storage)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    n0 = _cabi.launch_count()
    t0 = time.perf_counter()
    state = adcc.run_adc(matrix, output=None, **kw)
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t0
    mv = state.solver_state.timer.total("projection")
    r = dict(case=name, method=method, storage=matrix.doubles_storage, converged=bool(state.converged), n_iter=state.n_iter,
             n_applies=state.n_applies, build_s=round(t_build, 3), solve_s=round(t_solve, 3),
             launches=_cabi.launch_count() - n0, mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 2),
             e0=float(state.excitation_energy[0]))
    print(json.dumps(r), flush=True)
    if os.environ.get("SHOW_TIMERS"):
        print(state.solver_state.timer.describe())
        print(matrix.timer.describe())
    res.append(r)
    del matrix, ref, state
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
