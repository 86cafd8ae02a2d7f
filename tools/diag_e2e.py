import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 62244000
h = torch.randn(n, dtype=torch.float64).pin_memory()
out = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    if rank == 0: d.copy_(h, non_blocking=True)
def h2d_new():
    if rank == 0:
        x = h.to("cuda", non_blocking=True)
def bc(): dist.broadcast(d, 0)
def d2h():
    if rank == 0: out.copy_(d, non_blocking=True)
res = dict(rank=rank, h2d=t(h2d), h2d_new=t(h2d_new), bcast=t(bc), d2h=t(d2h))
print(res, flush=True)
dist.barrier(); dist.destroy_process_group()
