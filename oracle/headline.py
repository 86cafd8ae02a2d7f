"""TEST INFRASTRUCTURE (oracle): the synthetic ADC(2)-x matvec of BASELINE configs[4] at FULL size
on the host cores - the CPU arm of bench.py and the whole-vector checker of the CUDA path.

Restates, on canonical index pairs (i<j, a<b, c<d - libtensor likewise evaluates symmetry-unique
blocks only), the reference equations
    adcc/adc_pp/matrix.py:353-375  ph_ph_2      (i1 / i2: :933-942, t2: adcc/LazyMp.py:37-54)
    adcc/adc_pp/matrix.py:270-276  ph_pphh_1
    adcc/adc_pp/matrix.py:288-294  pphh_ph_1
    adcc/adc_pp/matrix.py:127-133, 234-246  pphh_pphh_0 / _1
summed as adcc/AdcMatrix.py:390-396 does.  Inputs come from the counter-based definition
adcc_b200/synthetic.py::synthetic_eri_value (restated in C in oracle/csrc/headline_kernels.c so that
50.9 GB of <ab||cd> can be produced in seconds); the contractions are NumPy -> OpenBLAS DGEMMs on
all host threads, the index bookkeeping between the pair store and the GEMM operands is OpenMP C.

Parity: pinned by tests/test_oracle_headline_cpu.py against oracle.OracleAdcMatrix (the dense
restatement that reproduces the reference's known answers) on small shapes, bit-for-tolerance, and
directly against the reference's OWN matvec on the same counter-based integrals at nocc/nvirt = 6/10
and 8/20 (tests/golden/reference_adcc_synthetic_workload.npz,
tests/test_reference_goldens_cpu.py::test_oracles_on_synthetic_workload_vs_reference_run).
Nothing under adcc_b200/ imports this module.
"""
import ctypes
import os
import subprocess
import time

import numpy as np

from adcc_b200.synthetic import synthetic_orbital_energies

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "csrc", "headline_kernels.c")
_LIB = os.path.join(HERE, "_build", "liboracle_headline.so")
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp the C helper into oracle/_build/ (called by __graft_entry__.build())."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, _SRC], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp, i64, u64, dbl, it = (ctypes.c_void_p, ctypes.c_longlong, ctypes.c_uint64, ctypes.c_double,
                                 ctypes.c_int)
        L.orc_eri.restype = dbl
        L.orc_eri.argtypes = [u64, it, it, it, it, it, dbl]
        L.orc_fill_vvvv_pairs.argtypes = [it, u64, dbl, i64, i64, dp]
        L.orc_fill_dense.argtypes = [it, it, it, it, it, u64, dbl, dp]
        L.orc_fill_oooo_pairs.argtypes = [it, u64, dbl, dp]
        L.orc_fill_ovov_jbkc.argtypes = [it, it, u64, dbl, dp]
        L.orc_fill_ovvv_pairs.argtypes = [it, it, u64, dbl, dp]
        L.orc_fill_ooov_pairs.argtypes = [it, it, u64, dbl, dp]
        for name in ("orc_expand_iakc", "orc_expand_ij", "orc_expand_ab"):
            getattr(L, name).argtypes = [it, it, dp, dp]
        for name in ("orc_fold_iajb", "orc_fold_ij", "orc_fold_ab"):
            getattr(L, name).argtypes = [it, it, dp, dbl, dp]
        L.orc_set_threads.argtypes = [it]
        L.orc_get_threads.restype = it
        L.orc_set_threads(os.cpu_count() or 1)
        _lib = L
    return _lib


def _p(a):
    assert a.flags.c_contiguous and a.dtype == np.float64
    return a.ctypes.data_as(ctypes.c_void_p)


def n_pairs(n):
    return n * (n - 1) // 2


def host_bytes_needed(no, nv):
    """Peak host memory of build + matvec (operands + the o^2 v^2 temporaries of the M1 build)."""
    npo, npv, ov = n_pairs(no), n_pairs(nv), no * nv
    return 8 * (npv * npv + no * nv * npv + 8 * ov * ov + no * no * npv + nv * npo * nv + npo * nv * nv)


class OracleHeadline:
    """sigma = M u for the synthetic ADC(2)-x workload, vectors as (u1[no,nv], U[npo,npv])."""

    def __init__(self, no, nv, seed=20261019, scale=5e-4):
        L = lib()
        self.no, self.nv, self.seed, self.scale = no, nv, seed, scale
        self.npo, self.npv = n_pairs(no), n_pairs(nv)
        npo, npv, ov = self.npo, self.npv, no * nv
        t0 = time.perf_counter()
        e_o, e_v = synthetic_orbital_energies(no, nv)
        eo = np.concatenate((e_o[0::2], e_o[1::2]))
        ev = np.concatenate((e_v[0::2], e_v[1::2]))
        self.eo, self.ev = eo, ev
        i, j = np.triu_indices(no, 1)                       # row order differs from the pair index:
        order = np.argsort(j * (j - 1) // 2 + i)            # sort into IJ = j(j-1)/2 + i
        self.pi, self.pj = i[order], j[order]
        a, b = np.triu_indices(nv, 1)
        order = np.argsort(b * (b - 1) // 2 + a)
        self.pa, self.pb = a[order], b[order]
        # ---- pphh_pphh operands
        self.D0 = (ev[self.pa] + ev[self.pb])[None, :] - (eo[self.pi] + eo[self.pj])[:, None]
        self.W = np.empty((npv, npv))
        L.orc_fill_vvvv_pairs(nv, seed, scale, 0, npv, _p(self.W))
        self.O = np.empty((npo, npo))
        L.orc_fill_oooo_pairs(no, seed, scale, _p(self.O))
        self.Y = np.empty((ov, ov))
        L.orc_fill_ovov_jbkc(no, nv, seed, scale, _p(self.Y))
        # ---- couplings
        self.V = np.empty((no, nv, npv))
        L.orc_fill_ovvv_pairs(no, nv, seed, scale, _p(self.V))
        self.G = np.empty((npo, no, nv))
        L.orc_fill_ooov_pairs(no, nv, seed, scale, _p(self.G))
        self.Gi = np.ascontiguousarray(self.G.transpose(1, 0, 2)).reshape(no, npo * nv)   # [i,(JK,b)]
        # ---- ph_ph_2 folded into one (ov x ov) matrix
        self.M1 = self._build_m1()
        self.build_seconds = time.perf_counter() - t0

    def _build_m1(self):
        L = lib()
        no, nv, ov = self.no, self.nv, self.no * self.nv
        oovv = np.empty((no, no, nv, nv))
        L.orc_fill_dense(2, no, no, nv, nv, self.seed, self.scale, _p(oovv))
        den = (self.ev[None, None, :, None] + self.ev[None, None, None, :]
               - self.eo[:, None, None, None] - self.eo[None, :, None, None])
        t2 = oovv / den                                                       # LazyMp.py:37-54
        x = np.tensordot(t2, oovv, axes=([0, 1, 3], [0, 1, 3]))              # t2[ijac] <ij||bc>
        i1 = np.diag(self.ev) + 0.25 * (x + x.T)                             # matrix.py:933-936
        z = np.tensordot(t2, oovv, axes=([1, 2, 3], [1, 2, 3]))              # t2[ikab] <jk||ab>
        i2 = np.diag(self.eo) - 0.25 * (z + z.T)                             # matrix.py:939-942
        # term_t2_eri[ikac] = t2[ijab] <jk||bc> + <ij||ab> t2[jkbc]  (matrix.py:363-366), as [(i,a),(k,c)]
        t2m = np.ascontiguousarray(t2.transpose(0, 2, 1, 3)).reshape(ov, ov)
        del t2
        erm = np.ascontiguousarray(oovv.transpose(0, 2, 1, 3)).reshape(ov, ov)
        del oovv
        m1 = t2m @ erm
        m1 += erm @ t2m
        del t2m, erm
        m1 *= -0.5
        ovov = np.empty((no, nv, no, nv))
        L.orc_fill_dense(3, no, nv, no, nv, self.seed, self.scale, _p(ovov))
        m1 -= ovov.transpose(2, 1, 0, 3).reshape(ov, ov)                     # - <ja||ib> u[jb]
        del ovov
        m4 = m1.reshape(no, nv, no, nv)
        for i in range(no):                                                   # u[ib] i1[ab] - i2[ij] u[ja]
            m4[i, :, i, :] += i1
        for a in range(nv):
            m4[:, a, :, a] -= i2
        return m1

    # ------------------------------------------------------------------ the matvec
    def matvec(self, u1, U):
        L = lib()
        no, nv, npo, npv, ov = self.no, self.nv, self.npo, self.npv, self.no * self.nv
        u1 = np.ascontiguousarray(u1, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        # ---- pphh_pphh_1 (matrix.py:234-246; canonical orbitals: the Fock terms are D0 * u)
        S = self.D0 * U
        S += U @ self.W.T                       # 1/2 u[ijcd] <ab||cd> = sum over c<d
        S += self.O @ U                         # 1/2 <ij||kl> u[klab] = sum over k<l
        X = np.empty((ov, ov))
        L.orc_expand_iakc(no, nv, _p(U), _p(X))
        T = X @ self.Y.T                        # [(i,a),(j,b)] = u[ikac] <kb||jc>
        del X
        L.orc_fold_iajb(no, nv, _p(T), -1.0, _p(S))          # -4 A_ij A_ab
        del T
        # ---- pphh_ph_1 (matrix.py:288-294)
        c1 = np.matmul(u1[None, :, :], self.V)               # [j,i,AB] = u[ic] <jc||ab>
        L.orc_fold_ij(no, nv, _p(c1), -0.5, _p(S))           # + A_ij: 1/2 (c1[j,i] - c1[i,j])
        del c1
        c2 = np.matmul(self.G.transpose(0, 2, 1), u1)        # [IJ,a,b] = <ij||ka> u[kb]
        L.orc_fold_ab(no, nv, _p(np.ascontiguousarray(c2)), -0.5, _p(S))
        del c2
        # ---- ph_ph_2 + ph_pphh_1 (matrix.py:353-375, 270-276)
        s1 = (self.M1 @ u1.reshape(ov)).reshape(no, nv)
        Uh = np.empty((nv, npo * nv))
        L.orc_expand_ab(no, nv, _p(U), _p(Uh))
        s1 += 2.0 * (self.Gi @ Uh.T)                         # <jk||ib> u[jkab] = 2 sum over j<k
        del Uh
        row = np.zeros((no, no), dtype=np.int64)
        sgn = np.zeros((no, no))
        row[self.pi, self.pj] = row[self.pj, self.pi] = np.arange(npo)
        sgn[self.pi, self.pj], sgn[self.pj, self.pi] = 1.0, -1.0
        for j in range(no):                                   # u[ijbc] <ja||bc> = 2 sum over b<c
            Uj = U[row[:, j]] * sgn[:, j][:, None]            # [i, BC] = u[i,j,BC]
            s1 += 2.0 * (Uj @ self.V[j].T)
        return s1, S

    # flops of one matvec as executed here (for the GFLOP/s figure of the CPU arm)
    def flops(self):
        no, nv, npo, npv, ov = self.no, self.nv, self.npo, self.npv, self.no * self.nv
        return (2.0 * npo * npv * npv + 2.0 * npo * npo * npv + 2.0 * ov ** 3
                + 2 * 2.0 * no * no * nv * npv + 2 * 2.0 * npo * no * nv * nv + 2.0 * ov * ov)
