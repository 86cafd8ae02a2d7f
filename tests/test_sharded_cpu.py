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
    # ---- distributed vectors (virtual-pair column slabs): matvec, scalar products, host form
    from adcc_b200.packed import SlabDoubles, shared_host_vector
    lay = m._engine.layout
    assert m.vector_distribution == "slab"
    vs = adcc.AmplitudeVector(ph=vp.ph, pphh=SlabDoubles.from_full(vp.pphh, lay))
    ss = m.matvec(vs)
    assert isinstance(ss.pphh, SlabDoubles) and tuple(ss.pphh._data.shape) == (lay.npo, lay.wmax)
    full = ss.pphh.to_full()
    err = max(err, rel_err(ss.ph.to_ndarray(), osig["ph"]), rel_err(full.to_ndarray(), osig["pphh"]))
    assert float(ss.pphh._data[:, lay.w:].abs().sum()) == 0.0           # padding stays zero
    d_slab, d_full = vs @ ss, vp @ sig                                   # one all-reduce vs local
    same = same and abs(d_slab - d_full) < 1e-12 * abs(d_full)
    d_list = vs @ [vs, ss]
    same = same and abs(d_list[1] - d_full) < 1e-12 * abs(d_full) and abs(d_list[0] - (vp @ vp)) < 1e-12 * (vp @ vp)
    ref_el = sig.pphh[1, 0, 2, 1]                                        # element access across ranks
    assert abs(ss.pphh[1, 0, 2, 1] - ref_el) < 1e-13 * max(1.0, abs(ref_el), float(sig.pphh._data.abs().max()))
    sh_in, sh_out = shared_host_vector(m), shared_host_vector(m)
    if rank == 0:
        sh_in["ph"].copy_(vp.ph._data)
        sh_in["pphh"].copy_(vp.pphh._data)
    dist.barrier()
    m.matvec_host(sh_in, sh_out)
    same = same and torch.allclose(sh_out["ph"], sig.ph._data, rtol=0, atol=1e-13 * float(sig.ph._data.abs().max()))
    same = same and torch.allclose(sh_out["pphh"], sig.pphh._data, rtol=0, atol=1e-13 * float(sig.pphh._data.abs().max()))
    # Davidson on distributed vectors (default) and on replicated vectors: same states, same iterations
    st = adcc.run_adc(m, n_states=2, conv_tol=1e-7, output=None)
    assert isinstance(st.excitation_vector[0].pphh, SlabDoubles)
    mr = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed", shard=(rank, world), vectors="replicated")
    assert mr.vector_distribution == "replicated"
    sr = adcc.run_adc(mr, n_states=2, conv_tol=1e-7, output=None)
    same = same and st.n_iter == sr.n_iter and np.max(np.abs(st.excitation_energy - sr.excitation_energy)) < 1e-10
    if rank == 0:
        ost = oracle.run_adc(om, "adc2x", n_states=2, conv_tol=1e-7)
        np.save(out, np.array([err, float(same), np.max(np.abs(st.excitation_energy - ost.eigenvalues)),
                               float(st.converged)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,no,nv", [(2, 6, 10), (3, 6, 8), (3, 8, 14)])
def test_sharded_matvec_gloo(tmp_path, world, no, nv):
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() + world) % 2000
    mp.spawn(_worker, args=(world, port, no, nv, out), nprocs=world, join=True)
    err, same, de, conv = np.load(out)
    assert err < 1e-11
    assert same == 1.0
    assert conv == 1.0 and de < 1e-8


def _worker_methods(rank, world, port, no, nv, out):
    """Sharded ADC(2) / ADC(3) on distributed vectors: sigma vs the oracle; ADC(3) Davidson."""
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
    from adcc_b200.packed import PackedDoubles, SlabDoubles, synthetic_reference_state
    from adc_parity_common import random_vector, rel_err
    res = []
    for method in ("adc2", "adc3"):
        # every operand of the sharded ADC(3) build comes from slabs / generated exchange slabs:
        # no rank ever holds the complete pair-packed vvvv or ovvv operand
        ref = synthetic_reference_state(no, nv, seed=5, scale=0.02)
        m = adcc.AdcMatrix(method, ref, doubles_storage="packed", shard=(rank, world))
        om = oracle.OracleAdcMatrix(method, OracleSyntheticRef(no, nv, seed=5, scale=0.02))
        v, ov = random_vector(m, om, 3)
        vs = adcc.AmplitudeVector(ph=v.ph, pphh=SlabDoubles.from_full(PackedDoubles.from_dense(v.pphh),
                                                                       m._engine.layout))
        ss, osig = m.matvec(vs), om.matvec(ov)
        res.append(max(rel_err(ss.ph.to_ndarray(), osig["ph"]), rel_err(ss.pphh.to_full().to_ndarray(), osig["pphh"])))
        d = m.diagonal()
        res.append(rel_err(d.ph.to_ndarray(), om.diagonal()["ph"]))
        if method == "adc3":
            st = adcc.run_adc(m, n_states=2, conv_tol=1e-7, output=None)
            ost = oracle.run_adc(om, method, n_states=2, conv_tol=1e-7)
            res.append(float(np.max(np.abs(st.excitation_energy - ost.eigenvalues))))
            res.append(float(st.converged))
    # restricted molecule (H2O / STO-3G), sharded, singlets and triplets on distributed vectors:
    # spin projection + purification act on the gathered vector, each rank keeps its slab
    from oracle.hfdata import load_h2o_sto3g
    data = load_h2o_sto3g()
    mh = adcc.AdcMatrix("adc2x", adcc.ReferenceState(data), doubles_storage="packed", shard=(rank, world))
    for kw in (dict(n_singlets=3), dict(n_triplets=3)):
        st = adcc.run_adc(mh, conv_tol=1e-8, output=None, **kw)
        ost = oracle.run_adc(data, "adc2x", conv_tol=1e-8, **kw)
        res.append(float(np.max(np.abs(st.excitation_energy - ost.eigenvalues))))
        res.append(float(st.converged and st.n_iter == ost.n_iter))
    if rank == 0:
        np.save(out, np.array(res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,no,nv", [(2, 6, 10), (3, 6, 8)])
def test_sharded_adc2_adc3_gloo(tmp_path, world, no, nv):
    out = str(tmp_path / "res.npy")
    port = 31500 + (os.getpid() + world) % 2000
    mp.spawn(_worker_methods, args=(world, port, no, nv, out), nprocs=world, join=True)
    e2, d2, e3, d3, de, conv, des, oks, det, okt = np.load(out)
    assert max(e2, e3) < 1e-11 and max(d2, d3) < 1e-11
    assert conv == 1.0 and de < 1e-8
    assert des < 1e-8 and det < 1e-8 and oks == 1.0 and okt == 1.0
