"""Host-overhead diagnosis for the small molecular-shape configs: run_adc wall time with the matvec
launched eagerly vs replayed from a CUDA graph, plus a cProfile of the replayed run."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adcc_b200 as adcc
from adcc_b200 import _cabi
from adcc_b200.synthetic import named_case

CASES = {"h2o_ccpvdz": ("adc2", dict(n_singlets=3, conv_tol=1e-6)),
         "h2o_ccpvtz": ("adc2", dict(n_singlets=10)),
         "methyloxirane_ccpvdz": ("adc3", dict(n_singlets=8)),
         "ne_ccpvtz_cvs": ("cvs-adc2x", dict(n_singlets=4, conv_tol=1e-8)),
         "h2o_ccpvtz_cvs": ("cvs-adc2x", dict(n_singlets=7, conv_tol=1e-8))}
names = sys.argv[1:] or ["h2o_ccpvtz"]
out = []
for name in names:
    method, kw = CASES[name]
    data, core = named_case(name)
    ref = adcc.ReferenceState(data, core_orbitals=core or None)
    for graphs in ("0", "1", "1"):
        os.environ["ADCCB_GRAPHS"] = graphs
        matrix = adcc.AdcMatrix(method, ref)
        torch.cuda.synchronize()
        n0 = _cabi.launch_count()
        t0 = time.perf_counter()
        state = adcc.run_adc(matrix, output=None, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r = dict(case=name, method=method, graphs=graphs, solve_s=round(dt, 4), n_iter=state.n_iter,
                 n_applies=state.n_applies, host_launches=_cabi.launch_count() - n0,
                 converged=bool(state.converged), e0=float(state.excitation_energy[0]))
        print(json.dumps(r), flush=True)
        out.append(r)
    if os.environ.get("PROFILE"):
        matrix = adcc.AdcMatrix(method, ref)
        adcc.run_adc(matrix, output=None, **kw)      # graphs captured
        pr = cProfile.Profile()
        pr.enable()
        adcc.run_adc(matrix, output=None, **kw)
        torch.cuda.synchronize()
        pr.disable()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
        print(buf.getvalue())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/diag_small.json", "w"), indent=1)
