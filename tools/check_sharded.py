"""Sharded vs single-GPU sigma on the same device-generated inputs (run under torchrun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import adcc_b200 as adcc
from adcc_b200.packed import PackedDoubles, synthetic_reference_state
no, nv = int(sys.argv[1]), int(sys.argv[2])
overlap = os.environ.get("ADCCB_NO_OVERLAP", "0") != "1"
ref = synthetic_reference_state(no, nv, seed=7, scale=5e-4)
m = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed", shard=(rank, world))
eng = m._engine
g = torch.Generator(device="cpu").manual_seed(3)
u1 = torch.randn(no, nv, generator=g, dtype=torch.float64).cuda()
u2 = torch.randn(eng.npo, eng.npv, generator=g, dtype=torch.float64).cuda()
v = adcc.AmplitudeVector(ph=adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], u1),
                         pphh=PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], u2))
outs = []
for rep in range(3):
    s = m.matvec(v)
    torch.cuda.synchronize()
    outs.append((s.ph._data.clone(), s.pphh._data.clone()))
same = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])
res = {"rank": rank, "run_to_run_identical": bool(same)}
if rank == 0:
    m1 = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed")
    s1 = m1.matvec(v)
    torch.cuda.synchronize()
    d2 = (s1.pphh._data - outs[0][1]).abs().max().item() / s1.pphh._data.abs().max().item()
    d1 = (s1.ph._data - outs[0][0]).abs().max().item() / s1.ph._data.abs().max().item()
    res.update(rel_err_ph=d1, rel_err_pphh=d2)
flat = torch.cat((outs[0][0].reshape(-1), outs[0][1].reshape(-1)))
ref0 = flat.clone()
dist.broadcast(ref0, 0)
res["identical_to_rank0"] = bool(torch.equal(ref0, flat))
print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
