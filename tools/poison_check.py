"""Uninitialised-read hunt: pre-fill the caching allocator's free blocks with NaN so that every
`empty` buffer starts poisoned; any output element a kernel forgets to write (or any read of
scratch before it is written) then shows up as NaN in the results."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adcc_b200 as adcc
from adcc_b200 import backend as be
from adcc_b200.packed import PackedDoubles, synthetic_reference_state

no, nv = int(sys.argv[1]), int(sys.argv[2])
gb = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0


def poison():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((int(gb * 2 ** 30 / 8 / 16),), float("nan"), dtype=torch.float64, device="cuda")
              for _ in range(16)]
    small = [torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
             for n in (64, 512, 4096, 65536, 2 ** 20) for _ in range(8)]
    be.workspace().fill_(float("nan"))
    torch.cuda.synchronize()
    del blocks, small


def nans(t):
    return int(torch.isnan(t).sum().item())


res = {}
poison()
ref = synthetic_reference_state(no, nv, seed=7, scale=5e-4)
m = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed")
eng = m._engine
torch.cuda.synchronize()
res["build"] = {k: nans(getattr(eng, k)) for k in ("M1", "V2", "G1", "G2", "D0", "W", "O", "Y")}
res["diag"] = nans(eng.diagonal_pphh._data)
g = torch.Generator(device="cpu").manual_seed(3)
u1 = torch.randn(no, nv, generator=g, dtype=torch.float64).cuda()
u2 = torch.randn(eng.npo, eng.npv, generator=g, dtype=torch.float64).cuda()
v = adcc.AmplitudeVector(ph=adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], u1),
                         pphh=PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], u2))
outs = []
for rep in range(3):
    poison()
    s = m.matvec(v)
    torch.cuda.synchronize()
    outs.append((s.ph._data.clone(), s.pphh._data.clone()))
    bad = torch.isnan(s.pphh._data).nonzero()
    info = {"nan_ph": nans(s.ph._data), "nan_pphh": int(bad.shape[0])}
    if bad.shape[0]:
        info["rows"] = [int(bad[:, 0].min()), int(bad[:, 0].max())]
        info["cols"] = [int(bad[:, 1].min()), int(bad[:, 1].max())]
    res[f"matvec{rep}"] = info
res["identical"] = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])
# dense engine on a smaller system (table interpreter, intermediates)
print(json.dumps(res), flush=True)
