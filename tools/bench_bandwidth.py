"""Achieved HBM bandwidth of the bandwidth-bound kernels at the headline vector size
(packed doubles vector 780 x 79800 = 0.498 GB), against MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from adcc_b200 import backend as be

PEAK = 6532.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
no, nv = 40, 400
npo, npv = no * (no - 1) // 2, nv * (nv - 1) // 2
n = npo * npv
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")   # 512 MB > L2


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


rows = []


def report(name, nbytes, fn):
    ms = timeit(fn)
    gbs = nbytes / ms * 1e-6
    rows.append((name, nbytes / 1e9, ms, gbs, gbs / PEAK))
    print(f"{name:44s} {nbytes/1e9:7.2f} GB {ms:8.3f} ms {gbs:8.0f} GB/s  {gbs/PEAK:5.2f} of measured copy peak", flush=True)


vecs = [torch.randn(npo, npv, device="cuda", dtype=torch.float64) for _ in range(21)]
out = torch.empty_like(vecs[0])
D = torch.rand(npo, npv, device="cuda", dtype=torch.float64) + 1.0
for k in (2, 10, 20):
    report(f"lincomb k={k} (residual / CGS / collapse)", 8 * n * (k + 1), lambda k=k: be.lincomb(np.ones(k), vecs[:k], out=out))
for m in (1, 10, 20):
    report(f"dot_many m={m} (p @ SS, subspace rows)", 8 * n * (m + 1), lambda m=m: be.dot_many(vecs[20], vecs[:m]))
report("div_shift r/(D-s) (Jacobi preconditioner)", 8 * n * 3, lambda: be.div_shift(vecs[0], D, 0.3))
report("mul D0*U (pphh_pphh_0, canonical)", 8 * n * 3, lambda: be.mul(vecs[0], D, out=out))
report("axpby", 8 * n * 3, lambda: be.axpby(0.5, vecs[0], 1.0, out))
report("scaled_copy (v / norm)", 8 * n * 2, lambda: be.scaled_copy(vecs[0], 0.5))
report("transpose U -> U^T (gather_tiled)", 8 * n * 2, lambda: be.permute(vecs[0], (1, 0)))
X = torch.empty(no * nv, no * nv, device="cuda", dtype=torch.float64)
report("expand_iakc U -> X[ia,kc]", 8 * (n + (no * nv) ** 2), lambda: be.packed_expand(0, vecs[0], no, nv, out=X))
Ue = torch.empty(no, no * npv, device="cuda", dtype=torch.float64)
report("expand_occ U -> Ue[i,(j,AB)]", 8 * (n + no * no * npv), lambda: be.packed_expand(1, vecs[0], no, nv, out=Ue))
Uh = torch.empty(nv, npo * nv, device="cuda", dtype=torch.float64)
report("expand_virt_t U -> Uh[a,(JK,b)]", 8 * (n + npo * nv * nv), lambda: be.packed_expand(2, vecs[0], no, nv, out=Uh))
pairs = [(i, j) for j in range(no) for i in range(j)]
off1 = be.int64_tensor([(i * nv * no + j) * nv for i, j in pairs])
off2 = be.int64_tensor([(j * nv * no + i) * nv for i, j in pairs])
report("fold_virt ovov: S += -(T - T' - T'' + T''')", 8 * ((no * nv) ** 2 + 2 * n),
       lambda: be.packed_fold_virt(out, nv, no * nv, X, off1, X, off2, alpha=-1.0))
report("fold_occ: S += A_ij(C)", 8 * (no * no * npv + 2 * n), lambda: be.packed_fold_occ(out, Ue, no, nv, alpha=0.5))
dov = torch.rand(no, nv, device="cuda", dtype=torch.float64)
report("packed_diagonal", 8 * n, lambda: be.packed_diagonal(no, nv, dov))
dense = torch.empty(no, no, nv, nv, device="cuda", dtype=torch.float64)
report("packed_unpack (to_ndarray path)", 8 * (n + no * no * nv * nv), lambda: be.packed_unpack(vecs[0], no, nv))
report("packed_pack (dense -> packed)", 8 * (n + no * no * nv * nv), lambda: be.packed_pack(dense, no, nv))
W = torch.empty(2048, npv, device="cuda", dtype=torch.float64)
report("synth_fill vvvv rows (hash generator)", 8 * W.numel(), lambda: be.synth_fill(W, "vvvv_packed", no, nv, 1, 0.02))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/bandwidth_kernels.txt", "w") as f:
    f.write(f"# achieved HBM bandwidth, L2 flushed between repetitions, peak = {PEAK} GB/s (MEASURED_PEAKS.json copy)\n")
    for name, gb, ms, gbs, frac in rows:
        f.write(f"{name:44s} {gb:7.2f} GB {ms:8.3f} ms {gbs:8.0f} GB/s  frac {frac:5.2f}\n")
