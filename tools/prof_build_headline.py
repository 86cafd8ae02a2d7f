"""Where the ADC(3) operand build at nocc=40 / nvirt=400 (packed_adc3.py, 308 TFLOP) spends its time:
cProfile with CUDA_LAUNCH_BLOCKING=1, so device time is attributed to the launching Python call."""
import cProfile
import io
import os
import pstats
import sys
import time

os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adcc_b200 as adcc
from adcc_b200.packed import synthetic_reference_state

no, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 400)
method = sys.argv[3] if len(sys.argv) > 3 else "adc3"
torch.zeros(1, device="cuda")
ref = synthetic_reference_state(no, nv, seed=20261019, scale=5e-4)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
matrix = adcc.AdcMatrix(method, ref, doubles_storage="packed")
torch.cuda.synchronize()
pr.disable()
print(f"build {time.perf_counter() - t0:.2f} s (launch-blocking), {matrix._engine.build_flops * 1e-12:.1f} TFLOP")
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(70)
print(buf.getvalue())
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(25)
print(buf.getvalue())
