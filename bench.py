#!/usr/bin/env python3
"""Headline benchmark: ADC(2)-x matrix-vector products per second on the synthetic
nocc=40 / nvirt=400 workload (BASELINE.json configs[4]; fits ONE B200 because the doubles
vectors and the vvvv block are stored antisymmetry-packed, 50.9 GB instead of 204.8 GB).

    python bench.py --gpus N --steps K --warmup W          # this framework (CUDA, sm_100a)
    python bench.py --impl reference ...                   # CPU arm: NumPy restatement (oracle)

One "step" = one sigma = M u through the public API (AdcMatrix.matvec).  Prints ONE JSON line.
  value      matvecs/s with the trial vector resident in HBM
  e2e        same through AdcMatrix.matvec with HOST (pinned) buffers: H2D of (u1, packed u2) and
             D2H of sigma inside the timed region
  roofline   dominant kernel (packed vvvv DMMA GEMM): executed flops / CUDA-event time against the
             DMMA.8x8x4 issue peak measured on this pool (profiles/r01_fp64_peak_microbench.txt;
             MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline  ONE complete matvec of the same workload by the oracle's full-size restatement
             (oracle/headline.py) on all host cores, fed with the same trial vector; its sigma is
             compared element by element with the GPU's (checks.full_sigma_rel_err_vs_cpu_oracle)
  --impl reference   times `steps` such complete CPU matvecs and prints the same line
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NO, NV = 40, 400
# ERI amplitude: uniform with standard deviation SCALE.  SURVEY 8(d) proposed 0.02, but the packed
# vvvv block is a 79800^2 random symmetric matrix whose spectral radius 2*SCALE*sqrt(79800) = 11.3
# would swamp the orbital-energy differences (0.5 .. 8) and leave Davidson chasing the edge of a
# semicircle spectrum; 5e-4 keeps the random part a perturbation (radius 0.28) so that the lowest
# states are singles-dominated and the solver converges.  Matvec cost does not depend on it.
SEED, SCALE = 20261019, 5e-4
FP64_DMMA_PEAK_TFLOPS = 37.17          # tools/microbench/fp64_peak.cu on this pool's B200
EXEC_FLOPS = 1.8448e13                 # executed flops of one packed ADC(2)-x matvec at nocc=40/nvirt=400
METRIC = "adc2x_matvecs_per_second"
UNIT = "matvec/s"


def dense_flops(no, nv):
    """As-written dense flops of one ADC(2)-x matvec (9.16e13 at 40/400)."""
    o, v = no, nv
    return (2.0 * o**2 * v**4 + 2.0 * o**3 * v**3 + 2.0 * o**4 * v**2     # vvvv, ovov, oooo terms
            + 3 * 2.0 * o**2 * v**3 + 3 * 2.0 * o**3 * v**2               # fvv/ovvv and foo/ooov terms
            + 2 * 2.0 * (o * v)**2)                                       # ph_ph GEMVs


def pair_flops(no, nv):
    """Flops of one ADC(2)-x matvec when the oooo / vvvv terms run over canonical index pairs (the
    oracle's unique_pairs mode; libtensor likewise works on symmetry-unique blocks): 1.9e13 at 40/400,
    essentially what the GPU path executes (1.84e13)."""
    o, v = no, nv
    npo, npv = o * (o - 1) // 2, v * (v - 1) // 2
    return (2.0 * npo * npv * npv + 2.0 * o**3 * v**3 + 2.0 * npo * npo * npv
            + 3 * 2.0 * o**2 * v**3 + 3 * 2.0 * o**3 * v**2 + 2 * 2.0 * (o * v)**2)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's full-size restatement (oracle/headline.py) on all host cores
# ------------------------------------------------------------------------------------------------
def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is to use every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


METHOD_LABEL = {"adc2": "ADC(2)", "adc2x": "ADC(2)-x", "adc3": "ADC(3)"}


def workload_config(no, nv, method="adc2x"):
    """`config` of the JSON line; identical in both arms when both run the same shape."""
    npv = nv * (nv - 1) // 2
    return {"workload": f"synthetic {METHOD_LABEL[method]} matvec nocc={no} nvirt={nv} (BASELINE configs[4]), "
                        "random antisymmetrised FP64 ERIs from a counter-based hash",
            "storage": f"antisymmetry-packed doubles + vvvv ({npv * npv * 8 / 1e9:.1f} GB instead of "
                       f"{8.0 * nv ** 4 / 1e9:.1f} GB dense)",
            "l2": "operands exceed the 126 MB L2 every step", "seed": SEED}


def _host_mem_available():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def cpu_shape(no, nv):
    """The requested shape if its operands fit the host memory, else the largest nvirt that does
    (the JSON line then names THAT shape and says so)."""
    from oracle.headline import host_bytes_needed
    if os.environ.get("ADCCB_BENCH_QUICK") == "1":          # contract tests: tiny, same code path
        return 6, 10
    avail = _host_mem_available()
    while nv > 40 and avail and host_bytes_needed(no, nv) * 1.15 > avail:
        nv -= 40
    return no, nv


def cpu_matvec_run(no, nv, n_warm, n_timed, u=None):
    """Build the full-size CPU restatement and time `n_timed` COMPLETE matvecs (wall clock).
    Returns (result dict, sigma of the last matvec)."""
    cores = _all_host_threads()
    from oracle.headline import OracleHeadline
    oh = OracleHeadline(no, nv, seed=SEED, scale=SCALE)
    if u is None:
        rng = np.random.default_rng(SEED + 1)
        u = (rng.standard_normal((no, nv)), rng.standard_normal((oh.npo, oh.npv)))
    sig = None
    for _ in range(n_warm):
        sig = oh.matvec(*u)
    times = []
    for _ in range(n_timed):
        t0 = time.perf_counter()
        sig = oh.matvec(*u)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    res = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": (f"{n_timed} complete matvec(s) at nocc={no} nvirt={nv} by the oracle's full-size "
                      f"restatement (oracle/headline.py: canonical index pairs like libtensor's "
                      f"symmetry-unique blocks, NumPy/OpenBLAS DGEMMs + OpenMP C index kernels, {cores} "
                      f"threads): {t:.2f} s per matvec = {oh.flops() / t * 1e-9:.0f} GFLOP/s executed; "
                      f"operand build {oh.build_seconds:.0f} s (not timed). libtensor itself is not "
                      "installable offline"),
           "seconds_per_matvec": t, "times": times, "build_s": oh.build_seconds}
    return res, sig


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    no, nv = cpu_shape(args.nocc, args.nvirt)
    res, _ = cpu_matvec_run(no, nv, args.warmup, max(1, args.steps))
    cfg = workload_config(no, nv)
    if (no, nv) != (args.nocc, args.nvirt):
        cfg["note"] = (f"host memory too small for nocc={args.nocc} nvirt={args.nvirt}: timed at "
                       f"nocc={no} nvirt={nv} instead (not extrapolated)")
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * res["seconds_per_matvec"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "build_s": res["build_s"], "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def sharded_vs_single(adcc, PackedDoubles, synthetic_reference_state, rank, world, no, nv):
    """N > 1: the complete sigma of the sharded engine against the single-GPU engine on the same
    device-generated inputs (nocc=40 / nvirt=240: 6.6 GB of packed vvvv, fits every rank), all ranks."""
    import torch
    import torch.distributed as dist
    ref = synthetic_reference_state(no, nv, seed=SEED + 7, scale=SCALE)
    ms = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed", shard=(rank, world))
    eng = ms._engine
    g = torch.Generator(device="cpu").manual_seed(SEED + 3)
    u1 = torch.randn(no, nv, generator=g, dtype=torch.float64).cuda()
    u2 = torch.randn(eng.npo, eng.npv, generator=g, dtype=torch.float64).cuda()
    v = adcc.AmplitudeVector(ph=adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], u1),
                             pphh=PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], u2))
    from adcc_b200.packed import SlabDoubles
    s_rep = ms.matvec(v)                      # complete vectors in, complete sigma out on every rank
    errs = {}
    if ms.vector_distribution == "slab":      # distributed vectors: slabs in, slabs out
        vs = adcc.AmplitudeVector(ph=v.ph, pphh=SlabDoubles.from_full(v.pphh, eng.layout))
        s_slab = ms.matvec(vs)
        s_sh = adcc.AmplitudeVector(ph=s_slab.ph, pphh=s_slab.pphh.to_full())
    else:
        s_sh = ms.matvec(v)                   # second call: both peer exchange buffers have been used
    m1 = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed")
    s_1 = m1.matvec(v)
    err = max(float((s[b]._data - s_1[b]._data).abs().max() / s_1[b]._data.abs().max())
              for b in ("ph", "pphh") for s in (s_sh, s_rep))
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    flat = torch.cat((s_sh.ph._data.reshape(-1), s_sh.pphh._data.reshape(-1)))
    ref0 = flat.clone()
    dist.broadcast(ref0, 0)
    same = torch.tensor([1.0 if torch.equal(ref0, flat) else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    del ms, m1, eng, s_sh, s_rep, s_1, v
    torch.cuda.empty_cache()
    return {"sharded_vs_single_rel": float(t.item()), "sharded_identical_on_all_ranks": bool(same.item() == 1.0),
            "sharded_check_shape": f"nocc={no} nvirt={nv}, complete sigma"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import adcc_b200 as adcc
    from adcc_b200 import _cabi, backend as be
    from adcc_b200.packed import PackedDoubles, synthetic_reference_state

    no, nv = args.nocc, args.nvirt
    sharded_check = None
    if world > 1 and not args.no_sharded_check:
        sharded_check = sharded_vs_single(adcc, PackedDoubles, synthetic_reference_state, rank, world,
                                          min(no, 40), min(nv, 240))
    t0 = time.perf_counter()
    ref = synthetic_reference_state(no, nv, seed=SEED, scale=SCALE)
    method = args.method
    if method != "adc2x":
        # the CPU restatement and the per-entry definition check are written for the ADC(2)-x blocks
        args.no_cpu_baseline, args.verify_samples = True, 0
    matrix = adcc.AdcMatrix(method, ref, doubles_storage="packed",
                            **({"shard": (rank, world)} if world > 1 else {}))
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    eng = matrix._engine
    npo, npv = eng.npo, eng.npv

    # trial vector: seeded, normalised, antisymmetric by construction of the packed store
    g = torch.Generator(device="cpu").manual_seed(SEED + 1)
    u1_host = torch.randn(no, nv, generator=g, dtype=torch.float64).pin_memory()
    u2_host = torch.randn(npo, npv, generator=g, dtype=torch.float64).pin_memory()
    nrm = float(torch.sqrt((u1_host**2).sum() + 4.0 * (u2_host**2).sum()))
    u1_host /= nrm
    u2_host /= nrm
    sig1_host = torch.empty_like(u1_host).pin_memory()
    sig2_host = torch.empty_like(u2_host).pin_memory()

    slab = world > 1 and matrix.vector_distribution == "slab"

    def to_vec(h1, h2):
        """Host (pinned) -> device-resident vector (outside every timed region).  N > 1: rank 0
        uploads, NCCL broadcasts, each rank keeps its virtual-pair slab (distributed vectors)."""
        if world == 1 or rank == 0:
            d1 = h1.to("cuda", non_blocking=True)
            d2 = h2.to("cuda", non_blocking=True)
        else:
            d1 = torch.empty(h1.shape, dtype=torch.float64, device="cuda")
            d2 = torch.empty(h2.shape, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.broadcast(d1, 0)
            dist.broadcast(d2, 0)
        full = PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], d2)
        if slab:
            from adcc_b200.packed import SlabDoubles
            full = SlabDoubles.from_full(full, eng.layout)
        return adcc.AmplitudeVector(ph=adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], d1), pphh=full)

    def full_doubles(vec):
        """Complete packed doubles array of a (possibly distributed) vector; collective if distributed."""
        return vec.pphh.to_full()._data if slab else vec.pphh._data

    def to_host(s):
        if rank == 0:                      # every rank holds the full sigma; rank 0 returns it
            sig1_host.copy_(s.ph._data, non_blocking=True)
            sig2_host.copy_(s.pphh._data, non_blocking=True)

    v = to_vec(u1_host, u2_host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        sig = matrix.matvec(v)
    eng.enable_kernel_timing(True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sig = matrix.matvec(v)
    e1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ktimes = eng.kernel_times_ms()
    eng.enable_kernel_timing(False)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps

    # ---- end to end: host buffers in, host buffers out -------------------------------------------
    if slab:
        # host vectors in pinned memory shared by the ranks of the box: every rank moves its own
        # slab over its own PCIe link
        u_host, s_host = adcc.shared_host_vector(matrix), adcc.shared_host_vector(matrix)
        if rank == 0:
            u_host["ph"].copy_(u1_host)
            u_host["pphh"].copy_(u2_host)
        sig1_host, sig2_host = s_host["ph"], s_host["pphh"]
        dist.barrier()
    else:
        u_host = {"ph": u1_host, "pphh": u2_host} if rank == 0 else None
        s_host = {"ph": sig1_host, "pphh": sig2_host} if rank == 0 else None

    def e2e_step():                           # public API: AdcMatrix.matvec_host (pipelined PCIe);
        matrix.matvec_host(u_host, s_host)    # N > 1: rank 0 owns the host buffers
    e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    ms_e2e /= args.steps
    h2d = 8 * (u1_host.numel() + u2_host.numel())
    d2h = 8 * (sig1_host.numel() + sig2_host.numel())

    # ---- invariants at full size (size-independent parity properties) ---------------------------
    checks = {}
    if rank == 0 or world > 1:
        g2 = torch.Generator(device="cpu").manual_seed(SEED + 2)
        w = to_vec(torch.randn(no, nv, generator=g2, dtype=torch.float64).pin_memory(),
                   torch.randn(npo, npv, generator=g2, dtype=torch.float64).pin_memory())
        mw = matrix.matvec(w)
        lhs, rhs = v @ mw, sig @ w
        checks["hermiticity_rel"] = abs(lhs - rhs) / max(abs(lhs), 1e-300)
        again = matrix.matvec(v)
        checks["run_to_run_identical"] = bool(torch.equal(again.pphh._data, sig.pphh._data)
                                              and torch.equal(again.ph._data, sig.ph._data))
        sig_full2, v_full2 = full_doubles(sig), full_doubles(v)
        if rank == 0:
            # the host buffers filled by the end-to-end path hold the same sigma
            ref2 = sig_full2.cpu()
            ref1 = sig.ph._data.cpu()
            checks["e2e_rel_diff_to_resident"] = float(max(
                (sig2_host - ref2).abs().max() / ref2.abs().max(),
                (sig1_host - ref1).abs().max() / ref1.abs().max()))
        if rank == 0 and args.verify_samples > 0:
            # sampled sigma entries recomputed on the CPU from the definition of the inputs
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from verify_headline import verify
            wrap2 = lambda d: PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], d)   # noqa: E731
            res = verify(matrix, adcc.AmplitudeVector(ph=v.ph, pphh=wrap2(v_full2)),
                         adcc.AmplitudeVector(ph=sig.ph, pphh=wrap2(sig_full2)),
                         n_pphh=args.verify_samples, n_ph=max(1, args.verify_samples // 4))
            checks.update({k: float(x) for k, x in res.items()})
        if sharded_check is not None:
            checks.update(sharded_check)

    # ---- time to converged states ----------------------------------------------------------------
    davidson = None
    if args.davidson:
        barrier()
        td0 = time.perf_counter()
        state = adcc.run_adc(matrix, n_states=args.n_states, kind="any", n_guesses=2 * args.n_states,
                             conv_tol=1e-6, output=None, max_iter=args.max_iter)
        barrier()
        t_dav = time.perf_counter() - td0
        if world > 1:
            t = torch.tensor([t_dav], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_dav = float(t.item())
        davidson = {"n_states": args.n_states, "conv_tol": 1e-6, "converged": bool(state.converged),
                    "n_iter": int(state.n_iter), "n_applies": int(state.n_applies),
                    "time_to_converged_s": t_dav,
                    "vectors": matrix.vector_distribution,
                    "excitation_energies": [float(e) for e in state.excitation_energy]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    flops_per_matvec = eng.flops_per_matvec
    eng_build = {"flops": eng.build_flops}
    vector_distribution = matrix.vector_distribution
    slab_exchange = ("none" if world == 1 else
                     "all-gather(u) + copy-engine pull of the partial-sigma slabs from peer memory under the "
                     "vvvv GEMM; vvvv slab needs no exchange" if slab and getattr(eng, "_pull", None) is not None else
                     "all-gather(u) + reduce-scatter(partial sigma); vvvv slab needs no exchange" if slab else
                     "GEMM epilogue -> peer memory (NVLink stores)" if eng._peer is not None
                     else "NCCL all-gather")
    f_ovov = eng.flops.get("pphh_pphh/ovov", 0.0)
    f_vvvv = eng.flops.get("pphh_pphh/vvvv", 0.0) / world
    t_vvvv = float(np.mean(ktimes["vvvv"])) if ktimes.get("vvvv") else None
    t_ovov = float(np.mean(ktimes["ovov"])) if ktimes.get("ovov") else None
    achieved = f_vvvv / (t_vvvv * 1e-3) * 1e-12 if t_vvvv else None
    roofline = {
        "bound": "tensor", "kernel": "dgemm_tn_tma_kernel (packed vvvv term, DMMA.8x8x4)",
        "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
        "frac": achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None,
        "peak_source": "measured DMMA.8x8x4 issue peak on this pool (profiles/r01_fp64_peak_microbench.txt); "
                       "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM reaches 36.2-36.5 here",
        "flops_per_launch": f_vvvv, "ms_per_launch": t_vvvv,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full capture
        # profiles/r02_vvvv_gemm.ncu-rep (1 GPU, nocc=40/nvirt=400): 89.44 GB + 0.53 GB per launch
        # (145.8 GB before the stream-K tail was shortened to the ragged round, 123.2 GB in round 1);
        # algorithmic: W 50.9 GB once + U 0.5 GB per round of tiles (29.5 rounds) + sigma RMW 1 GB =
        # 66.6 GB - the rest is the stream-K tail, whose spans share nothing through L2
        "traffic": 8.9967e10 if (world == 1 and no == NO and nv == NV) else None,
        "second_kernel": {"name": "dgemm_tn_tma_kernel (ovov term)", "flops_per_launch": f_ovov / world,
                          "ms_per_launch": t_ovov,
                          "achieved": f_ovov / world / (t_ovov * 1e-3) * 1e-12 if t_ovov else None},
        "share_of_step": (t_vvvv / ms_step) if t_vvvv else None,
        "section_ms": {k: float(np.mean(v)) for k, v in ktimes.items()},
    }
    if method == "adc2":
        # no vvvv / ovov GEMM at second order: the step is the two skinny ovvv GEMMs (M = nocc rows
        # against 2 x 10.2 GB of <ja||bc>) plus the M1 GEMV and the expand / fold / D0 kernels.  Both
        # lower bounds are stated; the larger one is the roofline of the step.
        hbm_peak = 6532.2
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        nbytes = 8.0 * (eng.V2.numel() + eng.M1.numel() / world + eng.G1.numel() + eng.G2.numel()
                        + 6.0 * npo * npv)
        step_flops = flops_per_matvec / world
        t_hbm_ms = nbytes / (hbm_peak * 1e9) * 1e3
        t_dmma_ms = step_flops / (FP64_DMMA_PEAK_TFLOPS * 1e12) * 1e3
        hbm_view = {"achieved": nbytes / (ms_step * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": t_hbm_ms / ms_step, "bytes_per_step": nbytes, "min_ms": t_hbm_ms,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"}
        dmma_view = {"achieved": step_flops / (ms_step * 1e-3) * 1e-12, "peak": FP64_DMMA_PEAK_TFLOPS,
                     "unit": "TFLOP/s", "frac": t_dmma_ms / ms_step, "flops_per_step": step_flops,
                     "min_ms": t_dmma_ms,
                     "peak_source": "measured DMMA.8x8x4 issue peak (profiles/r01_fp64_peak_microbench.txt)"}
        kernel = ("whole ADC(2) step: u1.V2 and Ue.V2^T GEMMs (DMMA.8x8x4) on the one ovvv layout, M1 GEMV, "
                  "expand / fold / D0 kernels")
        if t_dmma_ms >= t_hbm_ms:
            roofline = dict(dmma_view, bound="tensor", kernel=kernel, traffic=None, hbm_view=hbm_view)
        else:
            roofline = dict(hbm_view, bound="hbm", kernel=kernel, traffic=None, tensor_view=dmma_view)
        roofline["section_ms"] = {k: float(np.mean(v)) for k, v in ktimes.items()}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # one COMPLETE matvec of the same workload on the host cores (bounded: one step), fed with
        # the same trial vector; the whole sigma vector is compared with the GPU result
        cno, cnv = cpu_shape(no, nv)
        same = (cno, cnv) == (no, nv)
        u_np = (u1_host.numpy(), u2_host.numpy()) if same else None
        sig_ref = (sig.ph._data.cpu().numpy(), sig_full2.cpu().numpy()) if same else None
        del matrix, eng, sig, v
        torch.cuda.empty_cache()
        cpu, csig = cpu_matvec_run(cno, cnv, 0, 1, u=u_np)
        if same:
            checks["full_sigma_rel_err_vs_cpu_oracle"] = float(max(
                np.max(np.abs(csig[0] - sig_ref[0])) / np.max(np.abs(sig_ref[0])),
                np.max(np.abs(csig[1] - sig_ref[1])) / np.max(np.abs(sig_ref[1]))))
            checks["full_sigma_elements_compared"] = int(csig[0].size + csig[1].size)
    line = {
        "metric": METRIC.replace("adc2x", method), "value": 1000.0 / ms_step, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(no, nv, method),
        "parallelism": {"layout": f"virtual-pair slabs x{world}" if world > 1 else "single GPU",
                        "vectors": vector_distribution, "slab_exchange": slab_exchange},
        "executed_tflop_per_step": flops_per_matvec * 1e-12,
        "dense_tflop_per_step": dense_flops(no, nv) * 1e-12,
        "achieved_tflops_executed": flops_per_matvec / (ms_step * 1e-3) * 1e-12,
        "e2e": {"value": 1000.0 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
        "checks": checks,
        "build_s": t_build,
    }
    if eng_build["flops"]:
        line["build"] = {"seconds": t_build, "tflop": eng_build["flops"] * 1e-12,
                         "tflops_achieved": eng_build["flops"] * 1e-12 / t_build,
                         "note": "ADC(3) operands from pair-packed ERIs: o^2 v^4 slab GEMMs (t2sq.vvvv, ovov.(t2 t2), "
                                 "pi7), packed pi5 / pi3, ovov-type pi4 and the (ov)^3 terms of adc3_m11"}
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if davidson is not None:
        line["davidson"] = davidson
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _reserve_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its
    version banner on stdout): route fd 1 to stderr for the whole run and keep the real stdout for
    the result line."""
    global print
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    builtin_print = print

    def print(*a, **kw):      # noqa: A001 - module-level print() calls emit the result line
        kw.setdefault("file", real)
        kw.setdefault("flush", True)
        builtin_print(*a, **kw)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="adcc_b200", choices=["adcc_b200", "reference"])
    ap.add_argument("--method", default="adc2x", choices=["adc2", "adc2x", "adc3"],
                    help="ADC(2)-x is BASELINE's metric; adc2 / adc3 print the same line for the other methods")
    ap.add_argument("--nocc", type=int, default=NO)
    ap.add_argument("--nvirt", type=int, default=NV)
    ap.add_argument("--n-states", type=int, default=5)
    ap.add_argument("--max-iter", type=int, default=40)
    ap.add_argument("--no-sharded-check", action="store_true",
                    help="N > 1: skip the sharded-vs-single-GPU full sigma comparison at nocc=40 / nvirt=240")
    ap.add_argument("--verify-samples", type=int, default=64,
                    help="sigma entries re-derived on the CPU from the input definition (0 = off)")
    ap.add_argument("--no-davidson", dest="davidson", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__
        __graft_entry__.build()
        run_gpu(args)


if __name__ == "__main__":
    main()
