"""Where the one-shot build (ERI staging, MP, intermediates) spends its time: cProfile with
CUDA_LAUNCH_BLOCKING=1 so that device time is attributed to the launching call."""
import cProfile
import io
import os
import pstats
import sys

os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adcc_b200 as adcc
from adcc_b200.synthetic import named_case

name = sys.argv[1] if len(sys.argv) > 1 else "methyloxirane_ccpvdz"
method = sys.argv[2] if len(sys.argv) > 2 else "adc3"
data, core = named_case(name)
torch.zeros(1, device="cuda")
pr = cProfile.Profile()
pr.enable()
ref = adcc.ReferenceState(data, core_orbitals=core or None)
matrix = adcc.AdcMatrix(method, ref)
torch.cuda.synchronize()
pr.disable()
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(60)
print(buf.getvalue())
