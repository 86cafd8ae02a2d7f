"""Full-size parity of the packed ADC(2)-x matvec: sampled sigma entries recomputed on the CPU in
FP64 directly from the DEFINITION of the synthetic inputs (adcc_b200/synthetic.py hash) and the
working equations adcc/adc_pp/matrix.py:127-133, 234-246, 270-276, 288-294.
Independent of the device kernels' layouts: nothing here reads a staged GEMM operand or a
device-built intermediate - the singles-singles block is rebuilt on the host from the hash
(oovv -> t2 -> i1, i2, term_t2_eri rows; matrix.py:353-375, 933-942).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adcc_b200.synthetic import synthetic_eri_value, synthetic_orbital_energies  # noqa: E402


def _pairs(n):
    q = np.repeat(np.arange(n), np.arange(n))          # q for each pair, pairs ordered (p<q) by q then p
    p = np.concatenate([np.arange(k) for k in range(n)]) if n > 1 else np.zeros(0, dtype=int)
    return p.astype(int), q.astype(int)


def _dense_u(U, no, nv, i, k):
    """u[i,k,:,:] (nv x nv, antisymmetric) from the packed host array."""
    out = np.zeros((nv, nv))
    if i == k:
        return out
    row = U[max(i, k) * (max(i, k) - 1) // 2 + min(i, k)]
    a, b = _pairs(nv)
    out[a, b] = row
    out[b, a] = -row
    return out if i < k else -out


def sigma_pphh_entry(seed, scale, no, nv, u1, U, i, j, a, b):
    """sigma[i,j,a,b] of the pphh block rows for i<j, a<b (all four blocks feeding it)."""
    e_o, e_v = synthetic_orbital_energies(no, nv)
    eps_o = np.concatenate((e_o[0::2], e_o[1::2]))
    eps_v = np.concatenate((e_v[0::2], e_v[1::2]))
    IJ, AB = j * (j - 1) // 2 + i, b * (b - 1) // 2 + a
    pa, pb = _pairs(nv)
    po, qo = _pairs(no)
    val = lambda blk, *ix: synthetic_eri_value(seed, blk, *ix, scale)   # noqa: E731
    # pphh_pphh_0 (canonical orbitals)
    s = (eps_v[a] + eps_v[b] - eps_o[i] - eps_o[j]) * U[IJ, AB]
    # + 1/2 u[ijcd] <ab||cd> = sum_{c<d}
    s += float(np.dot(U[IJ], val("vvvv", a, b, pa, pb)))
    # + 1/2 <ij||kl> u[klab] = sum_{k<l}
    s += float(np.dot(val("oooo", i, j, po, qo), U[:, AB]))
    # -4 A_ij A_ab sum_kc u[ikac] <kb||jc>  =  -(T[ia,jb] - T[ja,ib] - T[ib,ja] + T[jb,ia])
    kk, cc = np.meshgrid(np.arange(no), np.arange(nv), indexing="ij")

    def T(i_, a_, j_, b_):
        tot = 0.0
        y = val("ovov", kk, b_, j_, cc)                  # <k b||j c>
        for k in range(no):
            if k == i_:
                continue
            tot += float(np.dot(_dense_u(U, no, nv, i_, k)[a_], y[k]))
        return tot
    s -= T(i, a, j, b) - T(j, a, i, b) - T(i, b, j, a) + T(j, b, i, a)
    # pphh_ph_1: A_ij(u[ic] <jc||ab>) - A_ab(<ij||ka> u[kb])
    c = np.arange(nv)
    s += 0.5 * (float(np.dot(u1[i], val("ovvv", j, c, a, b))) - float(np.dot(u1[j], val("ovvv", i, c, a, b))))
    k = np.arange(no)
    s -= 0.5 * (float(np.dot(val("ooov", i, j, k, a), u1[:, b])) - float(np.dot(val("ooov", i, j, k, b), u1[:, a])))
    return s


def sigma_ph_coupling_entry(seed, scale, no, nv, U, i, a):
    """ph_pphh_1 contribution to sigma[i,a]: <jk||ib> u[jkab] + u[ijbc] <ja||bc>."""
    val = lambda blk, *ix: synthetic_eri_value(seed, blk, *ix, scale)   # noqa: E731
    po, qo = _pairs(no)
    pa, pb = _pairs(nv)
    s = 0.0
    bb = np.arange(nv)
    # sum_{j<k} 2 <jk||ib> u[jk,a,b]
    g = val("ooov", po[:, None], qo[:, None], i, bb[None, :])            # [JK, b]
    ua = np.zeros((len(po), nv))
    lo, hi = np.minimum(a, bb), np.maximum(a, bb)
    ok = bb != a
    idx = hi[ok] * (hi[ok] - 1) // 2 + lo[ok]
    ua[:, ok] = U[:, idx] * np.where(a < bb[ok], 1.0, -1.0)[None, :]
    s += 2.0 * float(np.sum(g * ua))
    # sum_j sum_{b<c} 2 u[ij,bc] <ja||bc>
    for j in range(no):
        if j == i:
            continue
        row = U[max(i, j) * (max(i, j) - 1) // 2 + min(i, j)] * (1.0 if i < j else -1.0)
        s += 2.0 * float(np.dot(row, val("ovvv", j, a, pa, pb)))
    return s


class PhPhFromDefinition:
    """ph_ph_2 rows of the synthetic workload rebuilt on the host from the hash definition:
        sigma[ia] = u[ib] i1[ab] - i2[ij] u[ja] - <ja||ib> u[jb] - 1/2 term[ikac] u[kc]
    (adcc/adc_pp/matrix.py:353-375) with i1 / i2 (matrix.py:933-942), term_t2_eri (:363-366) and
    t2 = <ij||ab> / (e_a + e_b - e_i - e_j) (adcc/LazyMp.py:37-54).  Host memory: two o^2 v^2 arrays."""

    def __init__(self, seed, scale, no, nv):
        self.seed, self.scale, self.no, self.nv = seed, scale, no, nv
        e_o, e_v = synthetic_orbital_energies(no, nv)
        self.eps_o = np.concatenate((e_o[0::2], e_o[1::2]))
        self.eps_v = np.concatenate((e_v[0::2], e_v[1::2]))
        self.oovv = np.empty((no, no, nv, nv))
        if no * no * nv * nv > 10 ** 7:
            # the C restatement of the same hash (oracle/csrc, bit-identical to synthetic_eri_value:
            # tests/test_oracle_headline_cpu.py) on all host threads; NumPy needs 35 s at 40/400
            from oracle.headline import lib, _p
            lib().orc_fill_dense(2, no, no, nv, nv, seed, scale, _p(self.oovv))
        else:
            a, b = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
            for i in range(no):
                for j in range(no):
                    self.oovv[i, j] = synthetic_eri_value(seed, "oovv", i, j, a, b, scale)
        den = (self.eps_v[None, None, :, None] + self.eps_v[None, None, None, :]
               - self.eps_o[:, None, None, None] - self.eps_o[None, :, None, None])
        self.t2 = self.oovv / den
        x = np.tensordot(self.t2, self.oovv, axes=([0, 1, 3], [0, 1, 3]))          # [a,b]
        self.i1 = np.diag(self.eps_v) + 0.25 * (x + x.T)
        z = np.tensordot(self.t2, self.oovv, axes=([1, 2, 3], [1, 2, 3]))          # [i,j]
        self.i2 = np.diag(self.eps_o) - 0.25 * (z + z.T)

    def entry(self, u1, i, a):
        no, nv = self.no, self.nv
        s = float(np.dot(u1[i], self.i1[a])) - float(np.dot(self.i2[i], u1[:, a]))
        jj, bb = np.meshgrid(np.arange(no), np.arange(nv), indexing="ij")
        s -= float(np.sum(synthetic_eri_value(self.seed, "ovov", jj, a, i, bb, self.scale) * u1))
        # term[i,k,a,c] = sum_jb t2[ijab] oovv[jkbc] + oovv[ijab] t2[jkbc]
        term = (np.tensordot(self.t2[i, :, a, :], self.oovv, axes=([0, 1], [0, 2]))
                + np.tensordot(self.oovv[i, :, a, :], self.t2, axes=([0, 1], [0, 2])))     # [k,c]
        return s - 0.5 * float(np.sum(term * u1))


def verify(matrix, v, sigma, n_pphh=24, n_ph=4, seed_samples=11):
    """Returns max relative errors of sampled sigma entries against the CPU recomputation."""
    ref = matrix.reference_state
    syn = ref.synthetic
    no, nv, seed, scale = syn["no"], syn["nv"], syn["seed"], syn["scale"]
    u1 = v.ph.to_ndarray()
    U = v.pphh._data.cpu().numpy()
    S = sigma.pphh._data.cpu().numpy()
    s1 = sigma.ph.to_ndarray()
    rng = np.random.default_rng(seed_samples)
    scale_pphh = np.max(np.abs(S))
    err2 = 0.0
    for _ in range(n_pphh):
        j = int(rng.integers(1, no)); i = int(rng.integers(0, j))
        b = int(rng.integers(1, nv)); a = int(rng.integers(0, b))
        ref_val = sigma_pphh_entry(seed, scale, no, nv, u1, U, i, j, a, b)
        got = S[j * (j - 1) // 2 + i, b * (b - 1) // 2 + a]
        err2 = max(err2, abs(got - ref_val) / scale_pphh)
    # ph: coupling part and singles-singles part, both from the definition of the inputs
    phph = PhPhFromDefinition(seed, scale, no, nv) if n_ph > 0 else None
    scale_ph = np.max(np.abs(s1))
    err1 = 0.0
    for _ in range(n_ph):
        i, a = int(rng.integers(0, no)), int(rng.integers(0, nv))
        ref_val = sigma_ph_coupling_entry(seed, scale, no, nv, U, i, a)
        ref_val += phph.entry(u1, i, a)
        err1 = max(err1, abs(s1[i, a] - ref_val) / scale_ph)
    return {"sampled_rel_err_pphh": err2, "sampled_rel_err_ph": err1, "n_pphh": n_pphh, "n_ph": n_ph}
