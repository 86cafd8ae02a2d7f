"""Sharded ADC(2) / ADC(3) engines on distributed vectors vs the single-GPU packed engine on the same
device-generated inputs (run under torchrun): python tools/check_sharded_methods.py NOCC NVIRT"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import adcc_b200 as adcc
from adcc_b200.packed import PackedDoubles, SlabDoubles, synthetic_reference_state
no, nv = int(sys.argv[1]), int(sys.argv[2])
res = {"rank": rank}
for method in ("adc2", "adc3"):
    ref = synthetic_reference_state(no, nv, seed=7, scale=5e-4)
    m = adcc.AdcMatrix(method, ref, doubles_storage="packed", shard=(rank, world))
    eng = m._engine
    g = torch.Generator(device="cpu").manual_seed(3)
    u1 = torch.randn(no, nv, generator=g, dtype=torch.float64).cuda()
    u2 = torch.randn(eng.npo, eng.npv, generator=g, dtype=torch.float64).cuda()
    full = PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], u2)
    ph = adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], u1)
    vs = adcc.AmplitudeVector(ph=ph, pphh=SlabDoubles.from_full(full, eng.layout))
    outs = []
    for rep in range(2):
        ss = m.matvec(vs)
        outs.append((ss.ph._data.clone(), ss.pphh.to_full()._data.clone()))
    res[f"{method}_run_to_run_identical"] = bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))
    m1 = adcc.AdcMatrix(method, synthetic_reference_state(no, nv, seed=7, scale=5e-4), doubles_storage="packed")
    s1 = m1.matvec(adcc.AmplitudeVector(ph=ph, pphh=full))
    res[f"{method}_rel_err_ph"] = ((s1.ph._data - outs[0][0]).abs().max() / s1.ph._data.abs().max()).item()
    res[f"{method}_rel_err_pphh"] = ((s1.pphh._data - outs[0][1]).abs().max() / s1.pphh._data.abs().max()).item()
    dd = (m.diagonal().pphh.to_full()._data - m1.diagonal().pphh._data).abs().max().item()
    res[f"{method}_diag_err"] = max(dd, (m.diagonal().ph._data - m1.diagonal().ph._data).abs().max().item())
    del m, m1, eng
    torch.cuda.empty_cache()
print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
