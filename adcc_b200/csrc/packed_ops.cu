// Antisymmetry-packed doubles store: kernels that move between the packed layout
//   U[IJ, AB],  IJ = pair(i<j) = j(j-1)/2 + i,  AB = pair(a<b) = b(b-1)/2 + a
// and the GEMM-ready dense layouts the ADC(2)/ADC(2)-x matvec needs, plus the counter-based
// generator of the synthetic nocc=40/nvirt=400 workload.
//
// The doubles part of an ADC vector is antisymmetric in (i,j) and in (a,b)
// (adcc/guess/guess_zero.py:163-171, adcc/AdcMatrix.py:445-455); libtensor stores only the
// canonical blocks, here only the canonical ELEMENTS are stored (1/4 of the dense size), so the
// explicit doubles symmetrisation (AdcMatrix.py:449-455) is the identity on this store and
// every Davidson vector operation moves 4x fewer bytes.  All kernels are HBM-bound.
#include "common.cuh"
#include <algorithm>
#include <atomic>

namespace adccb {

__device__ __forceinline__ long long pair_index(int p, int q) {  // p < q
  return (long long)q * (q - 1) / 2 + p;
}
// inverse of pair_index: largest q with q(q-1)/2 <= P
__device__ __forceinline__ void pair_decode(long long P, int& p, int& q) {
  long long qq = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)P)) * 0.5);
  while (qq * (qq - 1) / 2 > P) --qq;
  while ((qq + 1) * qq / 2 <= P) ++qq;
  q = (int)qq;
  p = (int)(P - qq * (qq - 1) / 2);
}

// cheap exact decode for pair indices below 2^23 (single-precision sqrt + one correction step each way)
__device__ __forceinline__ void pair_decode32(int P, int& p, int& q) {
  int qq = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)P)) * 0.5f);
  while (qq * (qq - 1) / 2 > P) --qq;
  while ((qq + 1) * qq / 2 <= P) ++qq;
  q = qq;
  p = P - qq * (qq - 1) / 2;
}

// ---------------------------------------------------------------------------------------
// dense <-> packed
// ---------------------------------------------------------------------------------------
// out[i,j,a,b] = s_ij s_ab U[pair(ij), pair(ab)]
__global__ void unpack_dense_kernel(int no, int nv, const double* __restrict__ U,
                                    double* __restrict__ out) {
  const long long npv = (long long)nv * (nv - 1) / 2;
  const long long n = (long long)no * no * nv * nv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % nv), a = (int)((idx / nv) % nv);
    const int j = (int)((idx / ((long long)nv * nv)) % no), i = (int)(idx / ((long long)nv * nv * no));
    double v = 0.0;
    if (i != j && a != b) {
      const double s = ((i < j) == (a < b)) ? 1.0 : -1.0;
      v = s * U[pair_index(min(i, j), max(i, j)) * npv + pair_index(min(a, b), max(a, b))];
    }
    out[idx] = v;
  }
}

// U[IJ,AB] = 1/4 (T[ijab] - T[jiab] - T[ijba] + T[jiba])   (= doubles symmetrisation + pack)
__global__ void pack_dense_kernel(int no, int nv, const double* __restrict__ T,
                                  double* __restrict__ U) {
  const long long npv = (long long)nv * (nv - 1) / 2, npo = (long long)no * (no - 1) / 2;
  const long long n = npo * npv;
  const long long s0 = (long long)no * nv * nv, s1 = (long long)nv * nv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    int i, j, a, b;
    pair_decode(idx / npv, i, j);
    pair_decode(idx % npv, a, b);
    U[idx] = 0.25 * (T[i * s0 + j * s1 + (long long)a * nv + b] - T[j * s0 + i * s1 + (long long)a * nv + b] -
                     T[i * s0 + j * s1 + (long long)b * nv + a] + T[j * s0 + i * s1 + (long long)b * nv + a]);
  }
}

// ---------------------------------------------------------------------------------------
// GEMM-ready expansions of the packed trial vector
// ---------------------------------------------------------------------------------------
// Antisymmetric expansion of packed virtual-pair rows, 32x32 tiles through shared memory:
//   out[base_r + a*sA + c] = +s_r * row_r[pair(a,c)],  out[base_r + c*sA + a] = -s_r * row_r[pair(a,c)]
// for a < c, zero on the diagonal.  The packed row is read once, coalesced along a (pair(a,c) is
// contiguous in a for fixed c); both output triangles are written coalesced (c resp. a fastest).
//   mode 0: X[i,a,k,c] = u[i,k,a,c]  (ovov-term operand): r = (i,k), row = U[pair(i,k)], s = sign(k-i),
//           base = (i*nv*no + k)*nv, sA = no*nv; r with i == k is zero-filled
//   mode 2: Uh[a,JK,b] = u[JK,a,b]   : r = JK, row = U[JK], s = +1, base = JK*nv, sA = npo*nv
//   mode 0 with an occupied slab: r = (il, k), i = i_first + il; the output holds rows (il, a) only
__global__ void expand_antisym_kernel(int mode, int no, long long npo, int nv, int i_first,
                                      const double* __restrict__ U, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const long long npv = (long long)nv * (nv - 1) / 2;
  // blockIdx.x enumerates the tiles on and above the diagonal (ta <= tc); the tiles below it are
  // produced as mirrors, so no block is launched for them
  int tc = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while (tc * (tc + 1) / 2 > (int)blockIdx.x) --tc;
  while ((tc + 1) * (tc + 2) / 2 <= (int)blockIdx.x) ++tc;
  const int a0 = ((int)blockIdx.x - tc * (tc + 1) / 2) * 32, c0 = tc * 32;
  const long long r = blockIdx.z;
  const double* row;
  double sgn;
  long long base, sA;
  if (mode == 0) {
    const int il = (int)(r / no), k = (int)(r % no), i = i_first + il;
    sA = (long long)no * nv;
    base = ((long long)il * nv * no + k) * nv;
    if (i == k) { row = nullptr; sgn = 0.0; }
    else { row = U + pair_index(min(i, k), max(i, k)) * npv; sgn = (i < k) ? 1.0 : -1.0; }
  } else {
    row = U + r * npv; sgn = 1.0;
    sA = npo * nv;
    base = r * nv;
  }
  // tile[cc][aa] = sgn * row[pair(a0+aa, c0+cc)]  (0 unless a < c)
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const int c = c0 + cc, a = a0 + threadIdx.x;
    double v = 0.0;
    if (row != nullptr && a < c && c < nv) v = sgn * row[pair_index(a, c)];
    tile[cc][threadIdx.x] = v;
  }
  __syncthreads();
  // upper triangle out[a, c]: c fastest
  for (int aa = threadIdx.y; aa < 32; aa += blockDim.y) {
    const int a = a0 + aa, c = c0 + threadIdx.x;
    if (a < nv && c < nv && (a0 != c0 || a <= c)) out[base + (long long)a * sA + c] = tile[threadIdx.x][aa];
  }
  // lower triangle out[c, a] = -value: a fastest
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const int c = c0 + cc, a = a0 + threadIdx.x;
    if (a < nv && c < nv && a < c) out[base + (long long)c * sA + a] = -tile[cc][threadIdx.x];
  }
}

// Ue[i,jl,AB] = s_ij U[pair(ij),AB], j = j_first + jl < j_first + nj; one CTA per (i,jl), 16-byte
// accesses when the row length is even
__global__ void expand_occ_kernel(int no, int j_first, int nj, long long npv, const double* __restrict__ U,
                                  double* __restrict__ Ue) {
  const int i = blockIdx.x / nj, j = j_first + blockIdx.x % nj;
  double* dst = Ue + (long long)blockIdx.x * npv;
  const double s = (i == j) ? 0.0 : ((i < j) ? 1.0 : -1.0);
  const double* row = (i == j) ? U : U + pair_index(min(i, j), max(i, j)) * npv;
  if ((npv & 1) == 0) {
    const double2* src2 = reinterpret_cast<const double2*>(row);
    double2* dst2 = reinterpret_cast<double2*>(dst);
    for (long long e = threadIdx.x; e < npv / 2; e += blockDim.x) {
      double2 v = src2[e];
      v.x *= s; v.y *= s;
      dst2[e] = v;
    }
  } else {
    for (long long e = threadIdx.x; e < npv; e += blockDim.x) dst[e] = s * row[e];
  }
}

// ---------------------------------------------------------------------------------------
// epilogues that fold GEMM results back into the packed sigma vector
// ---------------------------------------------------------------------------------------
// S[r, pair(a,b)] += alpha * (F_r[a,b] - F_r[b,a]),  F_r[a,b] = P1[off1[r] + a*sa + b]
//                                                            - (P2 ? P2[off2[r] + a*sa + b] : 0)
// 32x32 tiles over (a,b) with a shared-memory transpose so that every global access is coalesced.
__global__ void pack_antisym_virt_kernel(int nv, long long sa, const double* __restrict__ P1,
                                         const long long* __restrict__ off1,
                                         const double* __restrict__ P2,
                                         const long long* __restrict__ off2,
                                         const long long* __restrict__ rows, double alpha,
                                         double* __restrict__ S) {
  __shared__ double tile[32][33];
  const long long npv = (long long)nv * (nv - 1) / 2;
  const long long r = blockIdx.z;
  // blockIdx.x enumerates the tiles with ta <= tb only (the others hold a > b everywhere)
  int tb = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while (tb * (tb + 1) / 2 > (int)blockIdx.x) --tb;
  while ((tb + 1) * (tb + 2) / 2 <= (int)blockIdx.x) ++tb;
  const int a0 = ((int)blockIdx.x - tb * (tb + 1) / 2) * 32, b0 = tb * 32;
  const double* F1 = P1 + off1[r];
  const double* F2 = P2 ? P2 + off2[r] : nullptr;
  // tile[bb][aa] = F[a0+aa, b0+bb] read coalesced along b
  for (int aa = threadIdx.y; aa < 32; aa += blockDim.y) {
    const int a = a0 + aa, b = b0 + threadIdx.x;
    double v = 0.0;
    if (a < nv && b < nv) {
      v = F1[(long long)a * sa + b];
      if (F2) v -= F2[(long long)a * sa + b];
    }
    tile[threadIdx.x][aa] = v;
  }
  __syncthreads();
  double* Srow = S + (rows ? rows[r] : r) * npv;
  for (int bb = threadIdx.y; bb < 32; bb += blockDim.y) {
    const int b = b0 + bb, a = a0 + threadIdx.x;
    if (a < b && b < nv) {
      double fba = F1[(long long)b * sa + a];
      if (F2) fba -= F2[(long long)b * sa + a];
      Srow[pair_index(a, b)] += alpha * (tile[bb][threadIdx.x] - fba);
    }
  }
}

// S[pair(ij), AB] += alpha * (C[i,j,AB] - C[j,i,AB]); one CTA per pair
__global__ void pack_antisym_occ_kernel(int no, long long npv, const double* __restrict__ C,
                                        double alpha, double* __restrict__ S) {
  int i, j;
  pair_decode(blockIdx.x, i, j);
  const double* c1 = C + ((long long)i * no + j) * npv;
  const double* c2 = C + ((long long)j * no + i) * npv;
  double* dst = S + (long long)blockIdx.x * npv;
  for (long long e = threadIdx.x; e < npv; e += blockDim.x) dst[e] += alpha * (c1[e] - c2[e]);
}

// slab form: C[i, jl, AB] holds occupied index j = j_begin + jl only.  pass 0 adds +alpha C[i,jl]
// to S[pair(i,j)] for i < j, pass 1 adds -alpha C[i,jl] to S[pair(j,i)] for i > j; inside one pass
// every S row is touched by one CTA only.
__global__ void pack_antisym_occ_slab_kernel(int no, int j_begin, int nj, long long npv, int pass,
                                             const double* __restrict__ C, double alpha,
                                             double* __restrict__ S) {
  const int i = blockIdx.x / nj, jl = blockIdx.x % nj, j = j_begin + jl;
  if (i == j || (pass == 0) != (i < j)) return;
  const double* c = C + (long long)blockIdx.x * npv;
  double* dst = S + pair_index(min(i, j), max(i, j)) * npv;
  const double a = (pass == 0) ? alpha : -alpha;
  for (long long e = threadIdx.x; e < npv; e += blockDim.x) dst[e] += a * c[e];
}

// D[IJ,AB] = 1/2 (dov[i,a] + dov[j,b] + dov[i,b] + dov[j,a]) + doo[i,j] + dvv[a,b]
// (packed form of adcc/adc_pp/matrix.py:109-124 (doo = dvv = 0, dov = e_a - e_i) and :212-231)
// grid (column chunks, rows): the occupied pair is decoded once per block, the virtual pair with a
// single-precision sqrt; no 64-bit division per element (the first version spent its time there:
// 0.15 of the copy peak)
__global__ void packed_diagonal_kernel(int no, int nv, const double* __restrict__ dov,
                                       const double* __restrict__ doo, const double* __restrict__ dvv,
                                       double* __restrict__ D) {
  const int npv = nv * (nv - 1) / 2, npo = no * (no - 1) / 2;
  for (int row = blockIdx.y; row < npo; row += gridDim.y) {
    int i, j;
    pair_decode32(row, i, j);
    const double* di = dov + (long long)i * nv;
    const double* dj = dov + (long long)j * nv;
    const double oo = doo ? doo[i * no + j] : 0.0;
    double* Drow = D + (long long)row * npv;
    for (int P = blockIdx.x * blockDim.x + threadIdx.x; P < npv; P += gridDim.x * blockDim.x) {
      int a, b;
      pair_decode32(P, a, b);
      double v = 0.5 * (di[a] + dj[b] + di[b] + dj[a]) + oo;
      if (dvv) v += dvv[(long long)a * nv + b];
      Drow[P] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// synthetic antisymmetrised integrals from a counter-based hash
//   <pq||rs> = sign * phi(seed, block, canonical(p,q,r,s)),  phi uniform in sqrt(3)*scale*[-1,1)
// layout codes select which tensor/layout is written (the value definition is shared with
// adcc_b200/synthetic.py::synthetic_eri_value, which the CPU oracle uses).
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

enum EriBlock { BLK_OOOO = 0, BLK_OOOV = 1, BLK_OOVV = 2, BLK_OVOV = 3, BLK_OVVV = 4, BLK_VVVV = 5 };

__host__ __device__ __forceinline__ double synthetic_eri(unsigned long long seed, int blk, int p,
                                                         int q, int r, int s, double scale) {
  double sign = 1.0;
  const bool bra_same = (blk == BLK_OOOO || blk == BLK_OOOV || blk == BLK_OOVV || blk == BLK_VVVV);
  const bool ket_same = (blk == BLK_OOOO || blk == BLK_OOVV || blk == BLK_OVVV || blk == BLK_VVVV);
  if (bra_same) {
    if (p == q) return 0.0;
    if (p > q) { int t = p; p = q; q = t; sign = -sign; }
  }
  if (ket_same) {
    if (r == s) return 0.0;
    if (r > s) { int t = r; r = s; s = t; sign = -sign; }
  }
  if (blk == BLK_OOOO || blk == BLK_VVVV || blk == BLK_OVOV) {
    if (p > r || (p == r && q > s)) { int t = p; p = r; r = t; t = q; q = s; s = t; }
  }
  const unsigned long long key = ((((unsigned long long)p * 1024ULL + q) * 1024ULL + r) * 1024ULL + s);
  const unsigned long long h = splitmix64(seed ^ splitmix64(key * 8ULL + (unsigned long long)blk));
  const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  return sign * (2.0 * u - 1.0) * 1.7320508075688772 * scale;
}

enum EriLayout {
  LAY_VVVV_PACKED = 0,   // W[AB,CD]
  LAY_OOOO_PACKED = 1,   // O[IJ,KL]
  LAY_OVOV_JBKC = 2,     // Y[(j,b),(k,c)] = <kb||jc>
  LAY_OVVV_JABC = 3,     // V1[(j,AB),c]  = <jc||ab>
  LAY_OVVV_AJBC = 4,     // V2[a,(j,BC)]  = <ja||bc>
  LAY_OOOV_IJKB = 5,     // G1[i,(JK,b)]  = <jk||ib>
  LAY_OOOV_IJAK = 6,     // G2[(IJ,a),k]  = <ij||ka>
  LAY_DENSE = 7,         // dense [p,q,r,s] of block `blk`
  LAY_VVVV_EXCH = 8,     // [al,b,c,d] = <a c||b d>, a = row_begin + al   (exchange-ordered slab)
  LAY_OVVV_EXCH = 9      // [al,c,j,d] = <j c||a d>, a = row_begin + al
};

// W[AB,CD] rows of the pair-packed vvvv block: grid (column chunks, rows), row pair decoded once per
// block, column pair with a single-precision sqrt (the generic kernel below pays two 64-bit divisions
// and two double-precision sqrt decodes per element: 0.12 of the write peak)
__global__ void synth_vvvv_rows_kernel(int nv, unsigned long long seed, double scale, long long row_begin,
                                       long long n_rows, double* __restrict__ out) {
  const int npv = nv * (nv - 1) / 2;
  for (long long r = blockIdx.y; r < n_rows; r += gridDim.y) {
    int a, b;
    pair_decode32((int)(row_begin + r), a, b);
    double* row = out + r * npv;
    for (int P = blockIdx.x * blockDim.x + threadIdx.x; P < npv; P += gridDim.x * blockDim.x) {
      int c, d;
      pair_decode32(P, c, d);
      row[P] = synthetic_eri(seed, BLK_VVVV, a, b, c, d, scale);
    }
  }
}

__global__ void synth_fill_kernel(int layout, int blk, int no, int nv, unsigned long long seed,
                                  double scale, long long n, long long row_begin, int j_begin,
                                  int nj, double* __restrict__ out) {
  const long long npv = (long long)nv * (nv - 1) / 2, npo = (long long)no * (no - 1) / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    int a, b, c, d, i, j, k, l;
    switch (layout) {
      case LAY_VVVV_PACKED: {
        pair_decode(row_begin + idx / npv, a, b);
        pair_decode(idx % npv, c, d);
        v = synthetic_eri(seed, BLK_VVVV, a, b, c, d, scale);
      } break;
      case LAY_OOOO_PACKED: {
        pair_decode(idx / npo, i, j);
        pair_decode(idx % npo, k, l);
        v = synthetic_eri(seed, BLK_OOOO, i, j, k, l, scale);
      } break;
      case LAY_OVOV_JBKC: {
        c = (int)(idx % nv); k = (int)((idx / nv) % no);
        b = (int)((idx / ((long long)nv * no)) % nv); j = j_begin + (int)(idx / ((long long)nv * no * nv));
        v = synthetic_eri(seed, BLK_OVOV, k, b, j, c, scale);
      } break;
      case LAY_OVVV_JABC: {
        c = (int)(idx % nv);
        pair_decode((idx / nv) % npv, a, b);
        j = j_begin + (int)(idx / ((long long)nv * npv));
        v = synthetic_eri(seed, BLK_OVVV, j, c, a, b, scale);
      } break;
      case LAY_OVVV_AJBC: {
        pair_decode(idx % npv, b, c);
        j = j_begin + (int)((idx / npv) % nj);
        a = (int)(idx / (npv * nj));
        v = synthetic_eri(seed, BLK_OVVV, j, a, b, c, scale);
      } break;
      case LAY_OOOV_IJKB: {
        b = (int)(idx % nv);
        pair_decode((idx / nv) % npo, j, k);
        i = (int)(idx / ((long long)nv * npo));
        v = synthetic_eri(seed, BLK_OOOV, j, k, i, b, scale);
      } break;
      case LAY_OOOV_IJAK: {
        k = (int)(idx % no); a = (int)((idx / no) % nv);
        pair_decode(idx / ((long long)no * nv), i, j);
        v = synthetic_eri(seed, BLK_OOOV, i, j, k, a, scale);
      } break;
      case LAY_VVVV_EXCH: {
        d = (int)(idx % nv); c = (int)((idx / nv) % nv); b = (int)((idx / ((long long)nv * nv)) % nv);
        a = (int)row_begin + (int)(idx / ((long long)nv * nv * nv));
        v = synthetic_eri(seed, BLK_VVVV, a, c, b, d, scale);
      } break;
      case LAY_OVVV_EXCH: {
        d = (int)(idx % nv); j = (int)((idx / nv) % no); c = (int)((idx / ((long long)nv * no)) % nv);
        a = (int)row_begin + (int)(idx / ((long long)nv * no * nv));
        v = synthetic_eri(seed, BLK_OVVV, j, c, a, d, scale);
      } break;
      default: {
        const int e3 = (blk == BLK_OOOO) ? no : nv;
        const int e2 = (blk == BLK_OOOO || blk == BLK_OOOV || blk == BLK_OVOV) ? no : nv;
        const int e1 = (blk == BLK_OVOV || blk == BLK_OVVV || blk == BLK_VVVV) ? nv : no;
        const int s = (int)(idx % e3), r = (int)((idx / e3) % e2);
        const int q = (int)((idx / ((long long)e3 * e2)) % e1), p = (int)(idx / ((long long)e3 * e2 * e1));
        v = synthetic_eri(seed, blk, p, q, r, s, scale);
      }
    }
    out[idx] = v;
  }
}


// out[l, pair(p,q), r] = in[l, p, q, r] for p < q  (staging of ERI blocks whose antisymmetric
// index pair is stored packed: ooov first pair, ovvv last pair)
__global__ void pair_select_kernel(long long L, int n, long long R, const double* __restrict__ in,
                                   double* __restrict__ out) {
  const long long np = (long long)n * (n - 1) / 2;
  const long long total = L * np * R;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx % R, P = (idx / R) % np, l = idx / (R * np);
    int p, q;
    pair_decode(P, p, q);
    out[idx] = in[((l * n + p) * n + q) * R + r];
  }
}

// Column slabs of a packed [nrows, npv] array: rank r owns columns [cut[r], cut[r+1]) and stores
// them as slab r of a [world, nrows, wmax] array (zero padded up to wmax columns).
//   dir 0: full -> slabs (input of the reduce-scatter that sums the ranks' partial sigma vectors)
//   dir 1: slabs -> full (output of the all-gather of a trial vector's slabs)
struct SlabCuts {
  int world;
  long long cut[17];
};
__global__ void slab_layout_kernel(int dir, long long nrows, long long npv, long long wmax, SlabCuts cuts,
                                   double* __restrict__ full, double* __restrict__ slabs) {
  const long long per = nrows * wmax;
  const long long n = per * cuts.world;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / per);
    const long long row = (idx % per) / wmax, c = idx % wmax;
    const long long col = cuts.cut[r] + c;
    const bool inside = col < cuts.cut[r + 1];
    if (dir == 0) slabs[idx] = inside ? full[row * npv + col] : 0.0;
    else if (inside) full[row * npv + col] = slabs[idx];
  }
}

// ---------------------------------------------------------------------------------------
// spin-block selection: compact copies of the rows / columns of one spin class
// ---------------------------------------------------------------------------------------
// mode 0: out[r, c] = in[rows[r] * ld + cols[c]]                  (gather a block)
// mode 1: in[rows[r] * ld + cols[c]] += alpha * out[r, c]         (scatter-add a block back)
// rows == nullptr / cols == nullptr: identity.  <pq||rs> vanishes unless the spin projections of bra
// and ket agree, so for a restricted reference the pair-packed vvvv / oooo operands and the ovov
// operand are block diagonal once their composite indices are ordered by spin class
// (libtensor keeps this structure in its spin-block symmetry elements,
// libadcc_src/TensorImpl/as_lt_symmetry.cc:100-172); these kernels move the classes in and out of
// the dense GEMM operands.
// mode 2: in[rows[r] * ld + cols[c]]  = alpha * out[r, c]         (scatter a block back, no accumulation)
// One thread per column (its column index is read once, coalesced), blockIdx.y strides over the rows:
// no per-element integer division, contiguous accesses on the compact side.
__global__ void block_select_kernel(int mode, long long nr, long long nc, const long long* __restrict__ rows,
                                    const long long* __restrict__ cols, long long ld, double alpha,
                                    double* __restrict__ big, double* __restrict__ blk) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const long long col = cols ? cols[c] : c;
  for (long long r = blockIdx.y; r < nr; r += gridDim.y) {
    double* b = big + (rows ? rows[r] : r) * ld + col;
    double* k = blk + r * nc + c;
    if (mode == 0) *k = *b;
    else if (mode == 1) *b += alpha * (*k);
    else *b = alpha * (*k);
  }
}

// ---------------------------------------------------------------------------------------
// ADC(3) build / matvec helpers on pair-packed operands
// ---------------------------------------------------------------------------------------
// Exchange-ordered slab of the vvvv block from its pair-packed form W[AB,CD] = <ab||cd>:
//   out[al, b, c, d] = <a c||b d>,  a = a_begin + al
// (operand of the o^2 v^4 builds of adc3_m11: -t2sq[idjc] <ac||bd>, adcc/adc_pp/matrix.py:1033, and
// of <ac||bd> p0[cd] in adc3_i1, :953).  One CTA per (al, c, b): the CTAs that run together share
// the packed row pair(a,c) through L2; the output is written coalesced along d.
__global__ void vvvv_exchange_slab_kernel(int nv, int a_begin, const double* __restrict__ W,
                                          double* __restrict__ out) {
  const long long npv = (long long)nv * (nv - 1) / 2;
  const int b = blockIdx.x % nv, c = (blockIdx.x / nv) % nv, al = blockIdx.x / (nv * nv);
  const int a = a_begin + al;
  double* dst = out + (((long long)al * nv + b) * nv + c) * nv;
  if (a == c) {
    for (int d = threadIdx.x; d < nv; d += blockDim.x) dst[d] = 0.0;
    return;
  }
  const double s_ac = (a < c) ? 1.0 : -1.0;
  const double* row = W + pair_index(min(a, c), max(a, c)) * npv;
  for (int d = threadIdx.x; d < nv; d += blockDim.x) {
    double v = 0.0;
    if (d != b) v = ((b < d) ? s_ac : -s_ac) * row[pair_index(min(b, d), max(b, d))];
    dst[d] = v;
  }
}

// Exchange-ordered slab of the ovvv block from V2[c,(j,AD)] = <jc||ad> (a<d packed):
//   out[al, c, j, d] = <j c||a d>,  a = a_begin + al
// (operand of pi7 = t2[ijbd] <jc||ad>, adcc/GroundState.py:239, an o^2 v^4 build).  One CTA per
// (al, c, j); output coalesced along d.
__global__ void ovvv_exchange_slab_kernel(int no, int nv, int a_begin, int na, const double* __restrict__ V2,
                                          double* __restrict__ out) {
  // one CTA per packed row (c, j); thread d walks the slab's a: for d > a the reads
  // row[pair(a, d)] are contiguous in a (the first version put a outermost: every read its own
  // 32-byte sector, 0.08 of the copy peak), for d < a contiguous in d
  const long long npv = (long long)nv * (nv - 1) / 2;
  const int j = blockIdx.x % no, c = blockIdx.x / no;
  const double* row = V2 + ((long long)c * no + j) * npv;
  for (int d = threadIdx.x; d < nv; d += blockDim.x) {
    for (int al = 0; al < na; ++al) {
      const int a = a_begin + al;
      double v = 0.0;
      if (d != a) v = ((a < d) ? 1.0 : -1.0) * row[pair_index(min(a, d), max(a, d))];
      out[(((long long)al * nv + c) * no + j) * nv + d] = v;
    }
  }
}

// y_rho[a] = sum_c A_rho[a,c] x_rho[c] for rows rho of an array of pair-packed ANTISYMMETRIC
// matrices (A_rho[b,c] = V[rho*npv + pair(b,c)] for b < c), summed over a reduction index:
//   partial[ch, o, a] = sum_{r in chunk ch} y_{rho(o,r)}[a],  rho(o,r) = o*so + r*sr,
//   x_rho = X + (x_by_out ? o : r) * ldx.
// Serves every contraction of ovvv (layout V2[a',(j,BC)] = <ja'||bc>) or vvvv (W) over ONE member of
// the antisymmetric pair without unpacking it: the three-operand second-order coupling terms
// icab,jkcd,jkbd and ijac,kbcd,kd (adcc/adc_pp/matrix.py:498-499, 525, 529) and <ia||bc> p0[ic] of
// adc3_i1 (:951).  HBM-bound: every packed element is read once; warp w owns the columns
// c = w, w + nwarps, ..., keeps private accumulators for y[b] += A x[c] (b < c) and reduces
// y[c] -= sum_b A x[b] with shuffles, so the result does not depend on the launch geometry.
constexpr int kPairWarps = 8;
__global__ void __launch_bounds__(kPairWarps * 32)
pair_matvec_kernel(int nv, int n_red, int chunk, long long so, long long sr, const double* __restrict__ V,
                   const double* __restrict__ X, long long ldx, int x_by_out, double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* xs = sm;                      // [nv]
  double* y2 = sm + nv;                 // [nv]
  double* yp = sm + 2 * nv;             // [kPairWarps][nv]
  const long long npv = (long long)nv * (nv - 1) / 2;
  const int o = blockIdx.x, ch = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < nv * (kPairWarps + 1); e += blockDim.x) y2[e] = 0.0;
  const int r0 = ch * chunk, r1 = min(r0 + chunk, n_red);
  for (int r = r0; r < r1; ++r) {
    __syncthreads();                    // previous row done with xs
    const double* x = X + (long long)(x_by_out ? o : r) * ldx;
    for (int e = threadIdx.x; e < nv; e += blockDim.x) xs[e] = x[e];
    __syncthreads();
    const double* A = V + ((long long)o * so + (long long)r * sr) * npv;
    double* ypw = yp + warp * nv;
    for (int c = warp + 1; c < nv; c += kPairWarps) {      // column c holds pairs (b, c), b < c
      const double* col = A + (long long)c * (c - 1) / 2;
      const double xc = xs[c];
      double acc = 0.0;
      int b = lane;
      for (; b + 96 < c; b += 128) {
        const double a0 = col[b], a1 = col[b + 32], a2 = col[b + 64], a3 = col[b + 96];
        ypw[b] += a0 * xc; ypw[b + 32] += a1 * xc; ypw[b + 64] += a2 * xc; ypw[b + 96] += a3 * xc;
        acc += a0 * xs[b] + a1 * xs[b + 32] + a2 * xs[b + 64] + a3 * xs[b + 96];
      }
      for (; b < c; b += 32) {
        const double a0 = col[b];
        ypw[b] += a0 * xc;
        acc += a0 * xs[b];
      }
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) y2[c] -= acc;
    }
  }
  __syncthreads();
  double* dst = partial + ((long long)ch * gridDim.x + o) * nv;
  for (int a = threadIdx.x; a < nv; a += blockDim.x) {
    double s = y2[a];
#pragma unroll
    for (int w = 0; w < kPairWarps; ++w) s += yp[w * nv + a];
    dst[a] = s;
  }
}

// out[o*ldo + a] = beta * out[o*ldo + a] + alpha * sum_ch partial[ch, o, a]   (fixed order)
__global__ void chunk_sum_kernel(int n_out, int nv, int n_chunks, const double* __restrict__ partial,
                                 double alpha, double beta, double* __restrict__ out, long long ldo) {
  const long long n = (long long)n_out * nv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int ch = 0; ch < n_chunks; ++ch) s += partial[(long long)ch * n + idx];
    double* dst = out + (idx / nv) * ldo + idx % nv;
    *dst = (beta != 0.0 ? beta * (*dst) : 0.0) + alpha * s;
  }
}

// ---------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------
static inline int blocks_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, 64LL * sm_count()));
}

int slab_layout(cudaStream_t st, int dir, long long nrows, long long npv, int world, const long long* cut,
                long long wmax, double* full, double* slabs) {
  if (world < 1 || world > 16) return fail("slab_layout: 1 .. 16 ranks");
  SlabCuts c;
  c.world = world;
  for (int r = 0; r <= world; ++r) c.cut[r] = cut[r];
  for (int r = 0; r < world; ++r)
    if (cut[r + 1] < cut[r] || cut[r + 1] - cut[r] > wmax || cut[r + 1] > npv) return fail("slab_layout: bad cuts");
  const long long n = nrows * wmax * world;
  if (n == 0) return 0;
  slab_layout_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dir, nrows, npv, wmax, c, full, slabs);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int block_select(cudaStream_t st, int mode, long long nr, long long nc, const long long* rows,
                 const long long* cols, long long ld, double alpha, double* big, double* blk) {
  if (nr <= 0 || nc <= 0) return 0;
  const unsigned gx = (unsigned)((nc + 255) / 256);
  const unsigned gy = (unsigned)std::min<long long>(nr, std::max<long long>(1, (32LL * sm_count()) / gx));
  block_select_kernel<<<dim3(gx, gy), 256, 0, st>>>(mode, nr, nc, rows, cols, ld, alpha, big, blk);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int vvvv_exchange_slab(cudaStream_t st, int nv, int a_begin, int na, const double* W, double* out) {
  if (na <= 0 || nv < 2) return 0;
  if (a_begin < 0 || a_begin + na > nv) return fail("vvvv_exchange_slab: slab out of range");
  const long long blocks = (long long)na * nv * nv;
  if (blocks > 0x7fffffffLL) return fail("vvvv_exchange_slab: slab too large");
  vvvv_exchange_slab_kernel<<<(unsigned)blocks, 128, 0, st>>>(nv, a_begin, W, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int ovvv_exchange_slab(cudaStream_t st, int no, int nv, int a_begin, int na, const double* V2, double* out) {
  if (na <= 0 || nv < 2 || no < 1) return 0;
  if (a_begin < 0 || a_begin + na > nv) return fail("ovvv_exchange_slab: slab out of range");
  const long long blocks = (long long)nv * no;
  if (blocks > 0x7fffffffLL) return fail("ovvv_exchange_slab: too many rows");
  ovvv_exchange_slab_kernel<<<(unsigned)blocks, 128, 0, st>>>(no, nv, a_begin, na, V2, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int pair_matvec(cudaStream_t st, int nv, int n_out, int n_red, long long so, long long sr, const double* V,
                const double* X, long long ldx, int x_by_out, double alpha, double beta, double* out,
                long long ldo, double* ws, long long ws_bytes) {
  if (n_out <= 0 || nv < 1) return 0;
  if (n_red <= 0) {
    chunk_sum_kernel<<<blocks_for((long long)n_out * nv, 256), 256, 0, st>>>(n_out, nv, 0, ws, alpha, beta, out, ldo);
    ADCCB_LAUNCH_OK();
    count_launch();
    return 0;
  }
  // enough CTAs to fill the machine: split the reduction index into chunks
  const int want = 4 * sm_count();      // (8 x measured slower: 2.89 / 2.71 ms against 2.79 / 2.26 ms per 10.2 GB)
  int n_chunks = std::max(1, std::min(n_red, (want + n_out - 1) / n_out));
  int chunk = (n_red + n_chunks - 1) / n_chunks;
  n_chunks = (n_red + chunk - 1) / chunk;
  if ((long long)n_chunks * n_out * nv * 8 > ws_bytes) return fail("pair_matvec: workspace too small");
  if (n_chunks > 65535) return fail("pair_matvec: too many chunks");
  const int smem = (kPairWarps + 2) * nv * 8;
  if (smem > 200 * 1024) return fail("pair_matvec: nv too large for the shared-memory accumulators");
  static std::atomic<unsigned long long> attr_devices{0};
  const int dev = current_device();
  if (dev >= 64 || !((attr_devices.load(std::memory_order_acquire) >> dev) & 1ULL)) {
    ADCCB_CUDA_OK(cudaFuncSetAttribute(pair_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) attr_devices.fetch_or(1ULL << dev, std::memory_order_release);
  }
  pair_matvec_kernel<<<dim3(n_out, n_chunks), kPairWarps * 32, smem, st>>>(nv, n_red, chunk, so, sr, V, X, ldx,
                                                                          x_by_out, ws);
  ADCCB_LAUNCH_OK();
  chunk_sum_kernel<<<blocks_for((long long)n_out * nv, 256), 256, 0, st>>>(n_out, nv, n_chunks, ws, alpha, beta, out, ldo);
  ADCCB_LAUNCH_OK();
  count_launch(2);
  return 0;
}

int packed_unpack(cudaStream_t st, int no, int nv, const double* U, double* out) {
  const long long n = (long long)no * no * nv * nv;
  if (n == 0) return 0;
  unpack_dense_kernel<<<blocks_for(n, 256), 256, 0, st>>>(no, nv, U, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int packed_pack(cudaStream_t st, int no, int nv, const double* T, double* U) {
  const long long n = (long long)no * (no - 1) / 2 * ((long long)nv * (nv - 1) / 2);
  if (n == 0) return 0;
  pack_dense_kernel<<<blocks_for(n, 256), 256, 0, st>>>(no, nv, T, U);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int packed_expand_slab(cudaStream_t st, int which, int no, int nv, int first, int count, const double* U,
                       double* out) {
  // which 0: occupied slab i = first .. first+count-1 of X; which 1: slab of the second occupied
  // index j of Ue; which 2: `count` rows of the packed vector starting at U (first is ignored)
  const long long npv = (long long)nv * (nv - 1) / 2;
  if (npv == 0 || count <= 0) return 0;
  if (which != 2 && (first < 0 || first + count > no)) return fail("packed_expand: slab out of range");
  const int t = (nv + 31) / 32;
  if (which == 0) {
    // grid.z is limited to 65535 (i,k) rows: walk the occupied slab in chunks of whole i
    if (no > 65535) return fail("packed_expand: more than 65535 occupied spin orbitals");
    const int ci = std::max(1, 65535 / no);
    for (int i0 = 0; i0 < count; i0 += ci) {
      const int ni = std::min(ci, count - i0);
      expand_antisym_kernel<<<dim3(t * (t + 1) / 2, 1, (unsigned)(ni * no)), dim3(32, 8), 0, st>>>(
          0, no, 0, nv, first + i0, U, out + (long long)i0 * nv * no * nv);
      ADCCB_LAUNCH_OK();
    }
    count_launch();
    return 0;
  } else if (which == 1) {
    expand_occ_kernel<<<no * count, 512, 0, st>>>(no, first, count, npv, U, out);
  } else if (which == 2) {
    // grid.z is limited to 65535: walk the pair rows in chunks (row r of a chunk writes out + r*nv)
    for (long long r0 = 0; r0 < count; r0 += 65535) {
      const long long nr = std::min<long long>(65535, count - r0);
      expand_antisym_kernel<<<dim3(t * (t + 1) / 2, 1, (unsigned)nr), dim3(32, 8), 0, st>>>(2, no, count, nv, 0,
                                                                                       U + r0 * npv, out + r0 * nv);
      ADCCB_LAUNCH_OK();
    }
    count_launch();
    return 0;
  } else {
    return fail("packed_expand: unknown layout");
  }
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int packed_expand(cudaStream_t st, int which, int no, int nv, const double* U, double* out) {
  // no < 0 (which == 2 only): U is a slab of -no rows of the packed vector
  if (no < 0 && which != 2) return fail("packed_expand: row slabs are supported for layout 2 only");
  if (no < 0) return packed_expand_slab(st, 2, no, nv, 0, -no, U, out);
  if (which == 2) return packed_expand_slab(st, 2, no, nv, 0, no * (no - 1) / 2, U, out);
  return packed_expand_slab(st, which, no, nv, 0, no, U, out);
}

int packed_fold_virt(cudaStream_t st, long long nrows, int nv, long long sa, const double* P1,
                     const long long* off1, const double* P2, const long long* off2,
                     const long long* rows, double alpha, double* S) {
  if (nrows == 0 || nv < 2) return 0;
  const int t = (nv + 31) / 32;
  // grid.z is limited to 65535 rows per launch.  Without an explicit `rows` map row r is S row r,
  // so a chunk starting at r0 shifts S by r0 rows.
  const long long npv = (long long)nv * (nv - 1) / 2;
  for (long long r0 = 0; r0 < nrows; r0 += 65535) {
    const long long nr = std::min<long long>(65535, nrows - r0);
    dim3 grid(t * (t + 1) / 2, 1, (unsigned)nr);
    pack_antisym_virt_kernel<<<grid, dim3(32, 8), 0, st>>>(nv, sa, P1, off1 + r0, P2, off2 ? off2 + r0 : nullptr,
                                                           rows ? rows + r0 : nullptr, alpha,
                                                           rows ? S : S + r0 * npv);
    ADCCB_LAUNCH_OK();
  }
  count_launch();
  return 0;
}

int packed_fold_occ(cudaStream_t st, int no, int nv, int j_begin, int nj, const double* C,
                    double alpha, double* S) {
  const long long npv = (long long)nv * (nv - 1) / 2, npo = (long long)no * (no - 1) / 2;
  if (npv == 0 || npo == 0) return 0;
  if (j_begin == 0 && (nj <= 0 || nj == no)) {
    pack_antisym_occ_kernel<<<(unsigned)npo, 256, 0, st>>>(no, npv, C, alpha, S);
    ADCCB_LAUNCH_OK();
    count_launch();
    return 0;
  }
  if (j_begin < 0 || j_begin + nj > no) return fail("packed_fold_occ: slab out of range");
  for (int pass = 0; pass < 2; ++pass) {
    pack_antisym_occ_slab_kernel<<<(unsigned)(no * nj), 256, 0, st>>>(no, j_begin, nj, npv, pass, C, alpha, S);
    ADCCB_LAUNCH_OK();
  }
  count_launch(2);
  return 0;
}

int packed_diagonal(cudaStream_t st, int no, int nv, const double* dov, const double* doo,
                    const double* dvv, double* D) {
  const long long n = (long long)no * (no - 1) / 2 * ((long long)nv * (nv - 1) / 2);
  if (n == 0) return 0;
  if (nv >= 4096 || no >= 4096) return fail("packed_diagonal: extents must be < 4096");
  const long long npv_ = (long long)nv * (nv - 1) / 2, npo_ = (long long)no * (no - 1) / 2;
  const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((npv_ + 255) / 256, 64));
  const unsigned gy = (unsigned)std::min<long long>(npo_, 65535);
  packed_diagonal_kernel<<<dim3(gx, gy), 256, 0, st>>>(no, nv, dov, doo, dvv, D);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int synth_fill(cudaStream_t st, int layout, int blk, int no, int nv, unsigned long long seed,
               double scale, long long n, long long row_begin, int j_begin, int nj, double* out) {
  if (n <= 0) return 0;
  if (no >= 1024 || nv >= 1024) return fail("synth_fill: extents must be < 1024");
  if (nj <= 0) nj = no;
  const long long npv_ = (long long)nv * (nv - 1) / 2;
  if (layout == LAY_VVVV_PACKED && npv_ > 0 && n % npv_ == 0) {
    const long long n_rows = n / npv_;
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((npv_ + 255) / 256, 64));
    synth_vvvv_rows_kernel<<<dim3(gx, (unsigned)std::min<long long>(n_rows, 65535)), 256, 0, st>>>(
        nv, seed, scale, row_begin, n_rows, out);
    ADCCB_LAUNCH_OK();
    count_launch();
    return 0;
  }
  synth_fill_kernel<<<blocks_for(n, 256), 256, 0, st>>>(layout, blk, no, nv, seed, scale, n, row_begin,
                                                        j_begin, nj, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int pair_select(cudaStream_t st, long long L, int n, long long R, const double* in, double* out) {
  const long long total = L * ((long long)n * (n - 1) / 2) * R;
  if (total <= 0) return 0;
  pair_select_kernel<<<blocks_for(total, 256), 256, 0, st>>>(L, n, R, in, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace adccb
