"""GEMM-epilogue slab exchange through peer memory vs. plain GEMM + NCCL all-gather (torchrun, N>=2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from adcc_b200 import backend as be

M, NT, K = (int(x) for x in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cut = [min(NT, ((k * NT // world + 127) // 128) * 128) for k in range(world)] + [NT]
c0, c1 = cut[rank], cut[rank + 1]
g = torch.Generator(device="cpu").manual_seed(5)
A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda()
Bfull = torch.randn(NT, K, generator=g, dtype=torch.float64)
B = Bfull[c0:c1].contiguous().cuda()
plain = be.gemm(A, B)
wmax = max(cut[k + 1] - cut[k] for k in range(world))
slab = torch.zeros(M, wmax, dtype=torch.float64, device="cuda")
slab[:, :c1 - c0] = plain
gathered = torch.empty(world * M, wmax, dtype=torch.float64, device="cuda")
dist.all_gather_into_tensor(gathered, slab)
gathered = gathered.view(world, M, wmax)
ref = torch.cat([gathered[r][:, :cut[r + 1] - cut[r]] for r in range(world)], dim=1)

bufs = [be.PeerBuffer.create(M * NT) for _ in range(2)]
assert all(b is not None for b in bufs), "peer buffers unavailable"
token = torch.zeros(1, dtype=torch.float64, device="cuda")
res = {"rank": rank, "bad": []}
for rep in range(reps):
    peer = bufs[rep % 2]
    R = peer.local.view(M, NT)
    R.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    be.gemm_peers(A, B, R[:, c0:c1], [p for r, p in enumerate(peer.peer_ptrs) if r != rank],
                  col_offset=c0, ldc=NT)
    dist.all_reduce(token)
    out = R.clone()
    torch.cuda.synchronize()
    for r in range(world):
        blk = out[:, cut[r]:cut[r + 1]]
        want = ref[:, cut[r]:cut[r + 1]]
        nbad = int((~(blk == want)).sum().item())
        if nbad:
            idx = (~(blk == want)).nonzero()
            res["bad"].append({"rep": rep, "slab_of_rank": r, "n": nbad,
                               "nan": int(torch.isnan(blk).sum().item()),
                               "rows": [int(idx[:, 0].min()), int(idx[:, 0].max())],
                               "cols": [int(idx[:, 1].min()), int(idx[:, 1].max())]})
print(json.dumps(res), flush=True)
dist.barrier()
for b in bufs:
    b.close()
dist.destroy_process_group()
