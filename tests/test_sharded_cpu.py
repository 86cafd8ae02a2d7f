"""N>1 path: world_size-2 gloo run of the sharded packed matvec (per-rank slabs of the vvvv /
ovvv / ovov operands, one all-reduce of the packed sigma) on CPU tensors via the TEST-ONLY shim."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, no, nv, out):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)

    class MP:
        def setattr(self, obj, name, val): setattr(obj, name, val)
    import cpu_shim
    cpu_shim.install(MP())
    import adcc_b200 as adcc
    import oracle
    from oracle.synthetic import OracleSyntheticRef
    from adcc_b200.packed import PackedDoubles, synthetic_reference_state
    from adc_parity_common import random_vector, rel_err
    ref = synthetic_reference_state(no, nv, seed=5, scale=0.02)
    m = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed", shard=(rank, world))
    om = oracle.OracleAdcMatrix("adc2x", OracleSyntheticRef(no, nv, seed=5, scale=0.02))
    v, ov = random_vector(m, om, 3)
    vp = adcc.AmplitudeVector(ph=v.ph, pphh=PackedDoubles.from_dense(v.pphh))
    sig, osig = m.matvec(vp), om.matvec(ov)
    err = max(rel_err(sig.ph.to_ndarray(), osig["ph"]), rel_err(sig.pphh.to_ndarray(), osig["pphh"]))
    # every rank must hold the identical, complete sigma
    flat = torch.cat((sig.ph._data.reshape(-1), sig.pphh._data.reshape(-1)))
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # end-to-end form: rank 0 owns the host buffers, every rank calls (plain exchange on CPU)
    host_in = {"ph": vp.ph._data.clone(), "pphh": vp.pphh._data.clone()} if rank == 0 else None
    host_out = {"ph": torch.zeros_like(vp.ph._data), "pphh": torch.zeros_like(vp.pphh._data)} if rank == 0 else None
    m.matvec_host(host_in, host_out)
    if rank == 0:
        same = same and torch.equal(host_out["ph"], sig.ph._data) and torch.equal(host_out["pphh"], sig.pphh._data)
    # Davidson on the sharded matrix (replicated vectors)
    st = adcc.run_adc(m, n_states=2, conv_tol=1e-7, output=None)
    if rank == 0:
        ost = oracle.run_adc(om, "adc2x", n_states=2, conv_tol=1e-7)
        np.save(out, np.array([err, float(same), np.max(np.abs(st.excitation_energy - ost.eigenvalues)),
                               float(st.converged)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,no,nv", [(2, 6, 10), (3, 6, 8)])
def test_sharded_matvec_gloo(tmp_path, world, no, nv):
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() + world) % 2000
    mp.spawn(_worker, args=(world, port, no, nv, out), nprocs=world, join=True)
    err, same, de, conv = np.load(out)
    assert err < 1e-11
    assert same == 1.0
    assert conv == 1.0 and de < 1e-8
