"""Parity at BASELINE's full size (nocc=40, nvirt=400, one B200): size-independent properties and
sampled sigma entries recomputed on the CPU from the input definition (tools/verify_headline.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _setup(no, nv, scale):
    import torch
    import adcc_b200 as adcc
    from adcc_b200.packed import PackedDoubles, synthetic_reference_state
    ref = synthetic_reference_state(no, nv, seed=20261019, scale=scale)
    m = adcc.AdcMatrix("adc2x", ref, doubles_storage="packed")
    eng = m._engine
    g = torch.Generator(device="cpu").manual_seed(5)

    def vec():
        u1 = torch.randn(no, nv, generator=g, dtype=torch.float64).cuda()
        u2 = torch.randn(eng.npo, eng.npv, generator=g, dtype=torch.float64).cuda()
        return adcc.AmplitudeVector(ph=adcc.Tensor._wrap(ref.mospaces, ["o1", "v1"], u1),
                                    pphh=PackedDoubles._wrap(ref.mospaces, ["o1", "o1", "v1", "v1"], u2))
    return m, vec


def test_sampled_sigma_medium_size():
    from verify_headline import verify
    m, vec = _setup(12, 40, 0.02)
    v = vec()
    res = verify(m, v, m.matvec(v), n_pphh=40, n_ph=10)
    assert res["sampled_rel_err_pphh"] < 1e-11 and res["sampled_rel_err_ph"] < 1e-11


@pytest.mark.parametrize("no,nv,edge", [(34, 60, 112), (34, 60, 96), (12, 40, 112)])
def test_matvec_between_host_buffers(no, nv, edge):
    """AdcMatrix.matvec_host (PCIe transfers pipelined with three row slabs of the vvvv GEMM; small
    systems take the plain path) fills the pinned result buffers with the same sigma as matvec."""
    import torch
    m, vec = _setup(no, nv, 0.02)
    v = vec()
    sig = m.matvec(v)
    u = {"ph": v.ph._data.cpu().pin_memory(), "pphh": v.pphh._data.cpu().pin_memory()}
    out = {"ph": torch.full_like(u["ph"], float("nan")).pin_memory(),
           "pphh": torch.full_like(u["pphh"], float("nan")).pin_memory()}
    for _ in range(3):                       # repeated: buffers are recycled across streams
        out["pphh"].fill_(float("nan"))
        m._engine.matvec_host(u["ph"], u["pphh"], out["ph"], out["pphh"], edge_rows=edge)
        for k, ref in (("ph", sig.ph._data), ("pphh", sig.pphh._data)):
            ref = ref.cpu()
            assert float((out[k] - ref).abs().max() / ref.abs().max()) < 1e-13
    m.matvec_host(u, out)


def test_headline_size_invariants_and_samples():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs a 180 GB class GPU")
    from verify_headline import verify
    m, vec = _setup(40, 400, 0.02)      # scale of SURVEY 8(d): strong off-diagonal terms
    v, w = vec(), vec()
    sv, sw = m.matvec(v), m.matvec(w)
    lhs, rhs = v @ sw, sv @ w                                   # <v|Mw> = <Mv|w>
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    assert torch.equal(m.matvec(v).pphh._data, sv.pphh._data)   # run-to-run identical
    # linearity
    z = m.matvec(2.0 * v - 0.5 * w)
    comb = 2.0 * sv - 0.5 * sw
    for k in ("ph", "pphh"):
        d = (z[k]._data - comb[k]._data).abs().max().item() / comb[k]._data.abs().max().item()
        assert d < 1e-12
    res = verify(m, v, sv, n_pphh=12, n_ph=3)
    assert res["sampled_rel_err_pphh"] < 1e-11 and res["sampled_rel_err_ph"] < 1e-11
