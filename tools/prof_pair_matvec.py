"""The two pair-matvec shapes of the ADC(3) matvec at nocc=40 / nvirt=400 on a 10.2 GB pair-packed ovvv
operand, a few launches each (run under ncu -k regex:pair_matvec_kernel; CUDA-event times printed for
the plain run)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adcc_b200 import backend as be

no, nv = 40, 400
npv = nv * (nv - 1) // 2
V2 = torch.randn(nv, no * npv, device="cuda", dtype=torch.float64)
r1 = torch.randn(nv, nv, device="cuda", dtype=torch.float64)
u1 = torch.randn(no, nv, device="cuda", dtype=torch.float64)
s1 = torch.zeros(no, nv, device="cuda", dtype=torch.float64)
for rep in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    be.pair_matvec(V2, nv, n_out=no, n_red=nv, row_stride_out=1, row_stride_red=no, X=r1, x_by_out=False, out=s1, beta=1.0)
    e[1].record()
    be.pair_matvec(V2, nv, n_out=nv, n_red=no, row_stride_out=no, row_stride_red=1, X=u1, x_by_out=False)
    e[2].record()
    torch.cuda.synchronize()
    gb = V2.numel() * 8e-9
    print(f"rep {rep}: out-by-i {e[0].elapsed_time(e[1]):.3f} ms ({gb / e[0].elapsed_time(e[1]) * 1e3:.0f} GB/s)  "
          f"out-by-b {e[1].elapsed_time(e[2]):.3f} ms ({gb / e[1].elapsed_time(e[2]) * 1e3:.0f} GB/s)", flush=True)
