/* TEST INFRASTRUCTURE (oracle): C restatement, for the host cores, of the pieces of the full-size
 * synthetic ADC(2)-x matvec that NumPy cannot do at memory speed on many cores: the counter-based
 * input definition (adcc_b200/synthetic.py::synthetic_eri_value) and the index bookkeeping between
 * the canonical-pair doubles store  U[IJ,AB] (i<j, a<b)  and the operands of the BLAS calls of
 * oracle/headline.py.  The contractions themselves are OpenBLAS DGEMMs issued from NumPy.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * Equations: adcc/adc_pp/matrix.py:234-246 (pphh_pphh_1), :270-276 (ph_pphh_1), :288-294
 * (pphh_ph_1); antisymmetrise = 1/2 (T - P T), libadcc_src/Tensor.hh:233-244.
 *
 *   gcc -O3 -fopenmp -shared -fPIC -o liboracle_headline.so headline_kernels.c
 */
#include <stdint.h>
#include <stddef.h>

#include <omp.h>

typedef long long i64;

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm is to use every host core */
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_get_threads(void) { return omp_get_max_threads(); }

static inline i64 pidx(i64 p, i64 q) { return q * (q - 1) / 2 + p; } /* p < q */

static inline uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* <pq||rs> of block blk: 0 oooo, 1 ooov, 2 oovv, 3 ovov, 4 ovvv, 5 vvvv */
double orc_eri(uint64_t seed, int blk, int p, int q, int r, int s, double scale) {
  double sign = 1.0;
  const int bra = (blk == 0 || blk == 1 || blk == 2 || blk == 5);
  const int ket = (blk == 0 || blk == 2 || blk == 4 || blk == 5);
  int t;
  if (bra) {
    if (p == q) return 0.0;
    if (p > q) { t = p; p = q; q = t; sign = -sign; }
  }
  if (ket) {
    if (r == s) return 0.0;
    if (r > s) { t = r; r = s; s = t; sign = -sign; }
  }
  if (blk == 0 || blk == 5 || blk == 3) {
    if (p > r || (p == r && q > s)) { t = p; p = r; r = t; t = q; q = s; s = t; }
  }
  const uint64_t key = ((((uint64_t)p * 1024ULL + (uint64_t)q) * 1024ULL + (uint64_t)r) * 1024ULL + (uint64_t)s);
  const uint64_t h = mix(seed ^ mix(key * 8ULL + (uint64_t)blk));
  const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  return sign * (2.0 * u - 1.0) * 1.7320508075688772 * scale;
}

/* W[AB,CD] = <ab||cd>, rows row_begin .. row_begin + n_rows - 1 of the pair-packed vvvv block */
void orc_fill_vvvv_pairs(int nv, uint64_t seed, double scale, i64 row_begin, i64 n_rows, double* W) {
  const i64 npv = (i64)nv * (nv - 1) / 2;
#pragma omp parallel for schedule(dynamic, 16)
  for (i64 r = 0; r < n_rows; ++r) {
    /* decode the row pair by scanning b (cheap: once per row) */
    i64 P = row_begin + r;
    int b = 1;
    while ((i64)(b + 1) * b / 2 <= P) ++b;
    const int a = (int)(P - (i64)b * (b - 1) / 2);
    double* row = W + r * npv;
    i64 Q = 0;
    for (int d = 1; d < nv; ++d)
      for (int c = 0; c < d; ++c) row[Q++] = orc_eri(seed, 5, a, b, c, d, scale);
  }
}

/* dense block [p,q,r,s] with extents e0..e3 */
void orc_fill_dense(int blk, int e0, int e1, int e2, int e3, uint64_t seed, double scale, double* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int p = 0; p < e0; ++p)
    for (int q = 0; q < e1; ++q) {
      double* o = out + ((i64)p * e1 + q) * e2 * e3;
      for (int r = 0; r < e2; ++r)
        for (int s = 0; s < e3; ++s) o[(i64)r * e3 + s] = orc_eri(seed, blk, p, q, r, s, scale);
    }
}

/* O[IJ,KL] = <ij||kl> */
void orc_fill_oooo_pairs(int no, uint64_t seed, double scale, double* O) {
  const i64 npo = (i64)no * (no - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int j = 1; j < no; ++j)
    for (int i = 0; i < j; ++i) {
      double* row = O + pidx(i, j) * npo;
      i64 Q = 0;
      for (int l = 1; l < no; ++l)
        for (int k = 0; k < l; ++k) row[Q++] = orc_eri(seed, 0, i, j, k, l, scale);
    }
}

/* Y[(j,b),(k,c)] = <kb||jc> */
void orc_fill_ovov_jbkc(int no, int nv, uint64_t seed, double scale, double* Y) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int j = 0; j < no; ++j)
    for (int b = 0; b < nv; ++b) {
      double* row = Y + ((i64)j * nv + b) * no * nv;
      for (int k = 0; k < no; ++k)
        for (int c = 0; c < nv; ++c) row[(i64)k * nv + c] = orc_eri(seed, 3, k, b, j, c, scale);
    }
}

/* V[j,c,AB] = <jc||ab>  (a<b); both coupling blocks read this one array */
void orc_fill_ovvv_pairs(int no, int nv, uint64_t seed, double scale, double* V) {
  const i64 npv = (i64)nv * (nv - 1) / 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int j = 0; j < no; ++j)
    for (int c = 0; c < nv; ++c) {
      double* row = V + ((i64)j * nv + c) * npv;
      i64 Q = 0;
      for (int b = 1; b < nv; ++b)
        for (int a = 0; a < b; ++a) row[Q++] = orc_eri(seed, 4, j, c, a, b, scale);
    }
}

/* G[IJ,k,a] = <ij||ka> */
void orc_fill_ooov_pairs(int no, int nv, uint64_t seed, double scale, double* G) {
#pragma omp parallel for schedule(static)
  for (int j = 1; j < no; ++j)
    for (int i = 0; i < j; ++i) {
      double* row = G + pidx(i, j) * no * nv;
      for (int k = 0; k < no; ++k)
        for (int a = 0; a < nv; ++a) row[(i64)k * nv + a] = orc_eri(seed, 1, i, j, k, a, scale);
    }
}

/* X[(i,a),(k,c)] = u[i,k,a,c] from the canonical-pair store */
void orc_expand_iakc(int no, int nv, const double* U, double* X) {
  const i64 npv = (i64)nv * (nv - 1) / 2, ld = (i64)no * nv;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < no; ++i)
    for (int a = 0; a < nv; ++a) {
      double* row = X + ((i64)i * nv + a) * ld;
      for (int k = 0; k < no; ++k) {
        double* dst = row + (i64)k * nv;
        if (k == i) { for (int c = 0; c < nv; ++c) dst[c] = 0.0; continue; }
        const double s = (i < k) ? 1.0 : -1.0;
        const double* u = U + pidx(i < k ? i : k, i < k ? k : i) * npv;
        for (int c = 0; c < a; ++c) dst[c] = -s * u[pidx(c, a)];
        dst[a] = 0.0;
        for (int c = a + 1; c < nv; ++c) dst[c] = s * u[pidx(a, c)];
      }
    }
}

/* Ue[i,j,AB] = u[i,j,AB] (sign of the occupied pair, zero for i == j) */
void orc_expand_ij(int no, int nv, const double* U, double* Ue) {
  const i64 npv = (i64)nv * (nv - 1) / 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < no; ++i)
    for (int j = 0; j < no; ++j) {
      double* dst = Ue + ((i64)i * no + j) * npv;
      if (i == j) { for (i64 e = 0; e < npv; ++e) dst[e] = 0.0; continue; }
      const double s = (i < j) ? 1.0 : -1.0;
      const double* u = U + pidx(i < j ? i : j, i < j ? j : i) * npv;
      for (i64 e = 0; e < npv; ++e) dst[e] = s * u[e];
    }
}

/* Uh[a,JK,b] = u[JK,a,b] */
void orc_expand_ab(int no, int nv, const double* U, double* Uh) {
  const i64 npv = (i64)nv * (nv - 1) / 2, npo = (i64)no * (no - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int a = 0; a < nv; ++a)
    for (i64 r = 0; r < npo; ++r) {
      const double* u = U + r * npv;
      double* dst = Uh + ((i64)a * npo + r) * nv;
      for (int b = 0; b < a; ++b) dst[b] = -u[pidx(b, a)];
      dst[a] = 0.0;
      for (int b = a + 1; b < nv; ++b) dst[b] = u[pidx(a, b)];
    }
}

/* S[IJ,AB] += alpha (T[ia,jb] - T[ja,ib] - T[ib,ja] + T[jb,ia]),  T[(i,a),(j,b)]  (= 4 A_ij A_ab) */
void orc_fold_iajb(int no, int nv, const double* T, double alpha, double* S) {
  const i64 npv = (i64)nv * (nv - 1) / 2, ld = (i64)no * nv, npo = (i64)no * (no - 1) / 2;
#pragma omp parallel for schedule(dynamic, 1)
  for (i64 r = 0; r < npo; ++r) {
    int j = 1;
    while ((i64)(j + 1) * j / 2 <= r) ++j;
    const int i = (int)(r - (i64)j * (j - 1) / 2);
    double* s = S + r * npv;
    for (int b = 1; b < nv; ++b)
      for (int a = 0; a < b; ++a)
        s[pidx(a, b)] += alpha * (T[((i64)i * nv + a) * ld + (i64)j * nv + b] - T[((i64)j * nv + a) * ld + (i64)i * nv + b] -
                                  T[((i64)i * nv + b) * ld + (i64)j * nv + a] + T[((i64)j * nv + b) * ld + (i64)i * nv + a]);
  }
}

/* S[IJ,AB] += alpha (C[i,j,AB] - C[j,i,AB]) */
void orc_fold_ij(int no, int nv, const double* C, double alpha, double* S) {
  const i64 npv = (i64)nv * (nv - 1) / 2;
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 1; j < no; ++j)
    for (int i = 0; i < j; ++i) {
      double* s = S + pidx(i, j) * npv;
      const double* c1 = C + ((i64)i * no + j) * npv;
      const double* c2 = C + ((i64)j * no + i) * npv;
      for (i64 e = 0; e < npv; ++e) s[e] += alpha * (c1[e] - c2[e]);
    }
}

/* S[IJ,(a,b)] += alpha (C[IJ,a,b] - C[IJ,b,a]) */
void orc_fold_ab(int no, int nv, const double* C, double alpha, double* S) {
  const i64 npv = (i64)nv * (nv - 1) / 2, npo = (i64)no * (no - 1) / 2;
#pragma omp parallel for schedule(static)
  for (i64 r = 0; r < npo; ++r) {
    double* s = S + r * npv;
    const double* c = C + r * nv * nv;
    for (int b = 1; b < nv; ++b)
      for (int a = 0; a < b; ++a) s[pidx(a, b)] += alpha * (c[(i64)a * nv + b] - c[(i64)b * nv + a]);
  }
}
