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
C5 = np.ones((5, 20))
report("lincomb_multi m=5 k=20 (5 residuals, one pass)", 8 * n * (20 + 5), lambda: be.lincomb_multi(C5, vecs[:20]))
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
# ---- round 2: ADC(3) build / matvec helpers, spin blocks, distributed-vector layout, fused correction step
del vecs[4:], X, Ue, Uh, dense, W
torch.cuda.empty_cache()
V2 = torch.randn(nv, no * npv, device="cuda", dtype=torch.float64)            # [a,(j,BC)], 10.2 GB
r1 = torch.randn(nv, nv, device="cuda", dtype=torch.float64)
u1 = torch.randn(no, nv, device="cuda", dtype=torch.float64)
s1 = torch.zeros(no, nv, device="cuda", dtype=torch.float64)
report("pair_matvec <ic||ab> r1[cb] (sum over c, out i)", 8 * V2.numel(),
       lambda: be.pair_matvec(V2, nv, n_out=no, n_red=nv, row_stride_out=1, row_stride_red=no, X=r1, x_by_out=False,
                              out=s1, beta=1.0))
report("pair_matvec r4[bc] = <kb||cd> u[kd] (sum over k)", 8 * V2.numel(),
       lambda: be.pair_matvec(V2, nv, n_out=nv, n_red=no, row_stride_out=no, row_stride_red=1, X=u1, x_by_out=False))
report("ovvv_exchange_slab (40 a: [a,c,j,d] from V2)", 8 * 40 * nv * no * nv,
       lambda: be.ovvv_exchange_slab(V2, no, nv, 100, 40))
del V2
torch.cuda.empty_cache()
Wn = 8000                                                                     # rows pair(a,c) up to a,c < 127
nvs = 120
npvs = nvs * (nvs - 1) // 2
Ws = torch.randn(npvs, npvs, device="cuda", dtype=torch.float64)
report(f"vvvv_exchange_slab (nv={nvs}, 60 a: [a,b,c,d] from packed W)", 8 * 60 * nvs ** 3,
       lambda: be.vvvv_exchange_slab(Ws, nvs, 30, 60))
del Ws
big = torch.empty(16, nv, nv, nv, device="cuda", dtype=torch.float64)
report("synth_fill exchange-ordered vvvv slab (16 a, hash)", 8 * big.numel(),
       lambda: be.synth_fill(big, "vvvv_exchange", no, nv, 1, 0.02, row_begin=100))
del big
cuts = [min(npv, ((k * npv // 8 + 127) // 128) * 128) for k in range(8)] + [npv]
wmax = max(cuts[k + 1] - cuts[k] for k in range(8))
slabs = torch.empty(8, npo, wmax, device="cuda", dtype=torch.float64)
report("slab_layout full -> 8 padded column slabs", 8 * 2 * n, lambda: be.slab_layout(0, vecs[0], slabs, cuts, wmax))
report("slab_layout 8 slabs -> full", 8 * 2 * n, lambda: be.slab_layout(1, vecs[1], slabs, cuts, wmax))
cols = torch.arange(0, npv, 3, device="cuda", dtype=torch.int64)
blk = torch.empty(npo, cols.numel(), device="cuda", dtype=torch.float64)
report("block_gather U[:, every 3rd column]", 8 * 2 * blk.numel(), lambda: be.block_gather(vecs[0], None, cols))
report("block_scatter_add S[:, cols] += blk", 8 * 3 * blk.numel(), lambda: be.block_scatter_add(vecs[1], blk, None, cols))
xd = torch.randn(24, 24, 140, 140, device="cuda", dtype=torch.float64)
dd = torch.rand(24, 24, 140, 140, device="cuda", dtype=torch.float64) + 1.0
report("precond_apply fused (dense 24x24x140x140, singlet)", 8 * 3 * xd.numel(),
       lambda: be.precond_apply(xd, dd, 0.3, [12, 12, 70, 70], sym_mode=1, spin_fac=1.0, purify=True))
report("correction step as shipped: div_shift, then gather kernel", 8 * 5 * xd.numel(),
       lambda: be.precond_apply(be.div_shift(xd, dd, 0.3), None, 0.0, [12, 12, 70, 70], sym_mode=1, spin_fac=1.0,
                                purify=True))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/bandwidth_kernels.txt", "w") as f:
    f.write(f"# achieved HBM bandwidth, L2 flushed between repetitions, peak = {PEAK} GB/s (MEASURED_PEAKS.json copy)\n")
    for name, gb, ms, gbs, frac in rows:
        f.write(f"{name:44s} {gb:7.2f} GB {ms:8.3f} ms {gbs:8.0f} GB/s  frac {frac:5.2f}\n")
