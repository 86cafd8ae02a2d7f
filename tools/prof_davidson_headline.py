"""Where the non-matvec time of the headline Davidson run goes (nocc=40 / nvirt=400 ADC(2)-x, 5 states):
the solver's own timers with a device synchronisation at every timer boundary (so device time is
attributed to the phase that launched it), plus the guess set-up and a cProfile of the host side."""
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
from adcc_b200 import timings
from adcc_b200.packed import synthetic_reference_state

no, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 400)
_stop = timings.Timer.stop
_restart = timings.Timer.restart


def stop_sync(self, task, now=None):
    torch.cuda.synchronize()
    return _stop(self, task, now)


def restart_sync(self, task):
    torch.cuda.synchronize()
    return _restart(self, task)


ref = synthetic_reference_state(no, nv, seed=20261019, scale=5e-4)
matrix = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed")
torch.cuda.synchronize()
out = {}
for sync in (False, True):
    timings.Timer.stop = stop_sync if sync else _stop
    timings.Timer.restart = restart_sync if sync else _restart
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if not sync:
        pr.enable()
    state = adcc.run_adc(matrix, n_states=5, kind="any", n_guesses=10, conv_tol=1e-6, output=None, max_iter=40)
    torch.cuda.synchronize()
    if not sync:
        pr.disable()
    dt = time.perf_counter() - t0
    tm = state.solver_state.timer
    out["sync" if sync else "async"] = dict(
        wall_s=dt, n_iter=int(state.n_iter), n_applies=int(state.n_applies),
        timers={t: [tm.total(t), len(tm.intervals.get(t, []))] for t in tm.tasks},
        matrix_timers={t: [matrix.timer.total(t), len(matrix.timer.intervals.get(t, []))] for t in matrix.timer.tasks})
    print(json.dumps(out["sync" if sync else "async"]), flush=True)
    if not sync:
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(40)
        print(buf.getvalue())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/prof_davidson_headline.json", "w"), indent=1)
