"""The full-size CPU restatement of the synthetic ADC(2)-x matvec (oracle/headline.py: canonical
pairs, C index bookkeeping, BLAS contractions) against the dense oracle that reproduces the
reference's known answers (oracle.OracleAdcMatrix), on shapes the dense form can hold."""
import numpy as np
import pytest

import oracle
from oracle.headline import OracleHeadline, lib
from oracle.synthetic import OracleSyntheticRef


def _pack(d, no, nv):
    i, j = np.triu_indices(no, 1)
    o = np.argsort(j * (j - 1) // 2 + i)
    i, j = i[o], j[o]
    a, b = np.triu_indices(nv, 1)
    o = np.argsort(b * (b - 1) // 2 + a)
    a, b = a[o], b[o]
    return d[i[:, None], j[:, None], a[None, :], b[None, :]]


@pytest.mark.parametrize("no,nv", [(4, 6), (6, 10), (8, 14)])
def test_headline_oracle_matches_dense_oracle(no, nv):
    seed, scale = 20261019, 0.02
    om = oracle.OracleAdcMatrix("adc2x", OracleSyntheticRef(no, nv, seed=seed, scale=scale))
    rng = np.random.default_rng(no * nv)
    v = {k: rng.standard_normal(s) for k, s in om.shapes.items()}
    v["pphh"] = om.symmetrise_doubles(v["pphh"])
    ref = om.matvec(v)
    oh = OracleHeadline(no, nv, seed=seed, scale=scale)
    s1, S = oh.matvec(v["ph"], _pack(v["pphh"], no, nv))
    assert np.max(np.abs(s1 - ref["ph"])) < 1e-12 * np.max(np.abs(ref["ph"]))
    want = _pack(ref["pphh"], no, nv)
    assert np.max(np.abs(S - want)) < 1e-12 * np.max(np.abs(want))


def test_c_hash_is_the_input_definition():
    from adcc_b200.synthetic import synthetic_eri_value
    L = lib()
    rng = np.random.default_rng(0)
    for blk, name in enumerate(("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")):
        for _ in range(200):
            p, q, r, s = (int(x) for x in rng.integers(0, 40, 4))
            want = float(synthetic_eri_value(20261019, name, p, q, r, s, 5e-4))
            assert L.orc_eri(20261019, blk, p, q, r, s, 5e-4) == want
