"""Multi-GPU sharded matvec over NCCL (needs >= 2 GPUs; skipped otherwise): bit-identical sigma on
all ranks, run-to-run identical, equal to the single-GPU engine."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("no,nv,pull", [(40, 240, "1"), (10, 38, "1"), (40, 240, "0")])    # nv = 38: odd pair count
def test_sharded_matches_single_gpu(no, nv, pull):
    """pull = "1": partial sigma exchanged by copy-engine pulls from peer-mapped memory under the vvvv
    GEMM (PackedAdcEngine._pull_exchange); "0": NCCL reduce-scatter."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29571",
           os.path.join(ROOT, "tools", "check_sharded.py"), str(no), str(nv)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, ADCCB_CE_EXCHANGE=pull))
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(rows) == world
    for r in rows:
        assert r["slab_exchange"] == ("copy-engine pull" if pull == "1" else "nccl reduce-scatter")
        assert r["run_to_run_identical"] and r["identical_to_rank0"]
        if r["rank"] == 0:
            assert r["rel_err_ph"] < 1e-11 and r["rel_err_pphh"] < 1e-11
            assert r["host_path_rel_diff"] < 1e-13
        assert r["slab_vs_replicated_rel"] < 1e-13 and r["slab_dot_rel"] < 1e-13
        assert r["slab_host_path_rel_diff"] < 1e-13


def test_sharded_adc2_adc3_match_single_gpu():
    """Sharded ADC(2) / ADC(3) (build dealt out over the ranks, distributed vectors) vs one GPU."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29573",
           os.path.join(ROOT, "tools", "check_sharded_methods.py"), "16", "60"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(rows) == world
    for r in rows:
        for method in ("adc2", "adc3"):
            assert r[f"{method}_run_to_run_identical"]
            assert r[f"{method}_rel_err_ph"] < 1e-11 and r[f"{method}_rel_err_pphh"] < 1e-11
            assert r[f"{method}_diag_err"] < 1e-11
