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
  const int a0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (a0 > c0) return;                          // tiles below the diagonal are produced as mirrors
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
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  if (a0 > b0 + 31) return;   // tile entirely below the diagonal (a > b everywhere)
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
__global__ void packed_diagonal_kernel(int no, int nv, const double* __restrict__ dov,
                                       const double* __restrict__ doo, const double* __restrict__ dvv,
                                       double* __restrict__ D) {
  const long long npv = (long long)nv * (nv - 1) / 2, npo = (long long)no * (no - 1) / 2;
  const long long n = npo * npv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    int i, j, a, b;
    pair_decode(idx / npv, i, j);
    pair_decode(idx % npv, a, b);
    double v = 0.5 * (dov[i * nv + a] + dov[j * nv + b] + dov[i * nv + b] + dov[j * nv + a]);
    if (doo) v += doo[i * no + j];
    if (dvv) v += dvv[(long long)a * nv + b];
    D[idx] = v;
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
  LAY_DENSE = 7          // dense [p,q,r,s] of block `blk`
};

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
    if ((long long)count * no > 65535) return fail("packed_expand: too many (i,k) rows");
    expand_antisym_kernel<<<dim3(t, t, count * no), dim3(32, 8), 0, st>>>(0, no, 0, nv, first, U, out);
  } else if (which == 1) {
    expand_occ_kernel<<<no * count, 512, 0, st>>>(no, first, count, npv, U, out);
  } else if (which == 2) {
    // grid.z is limited to 65535: walk the pair rows in chunks (row r of a chunk writes out + r*nv)
    for (long long r0 = 0; r0 < count; r0 += 65535) {
      const long long nr = std::min<long long>(65535, count - r0);
      expand_antisym_kernel<<<dim3(t, t, (unsigned)nr), dim3(32, 8), 0, st>>>(2, no, count, nv, 0, U + r0 * npv,
                                                                         out + r0 * nv);
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
    dim3 grid(t, t, (unsigned)nr);
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
  packed_diagonal_kernel<<<blocks_for(n, 256), 256, 0, st>>>(no, nv, dov, doo, dvv, D);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int synth_fill(cudaStream_t st, int layout, int blk, int no, int nv, unsigned long long seed,
               double scale, long long n, long long row_begin, int j_begin, int nj, double* out) {
  if (n <= 0) return 0;
  if (no >= 1024 || nv >= 1024) return fail("synth_fill: extents must be < 1024");
  if (nj <= 0) nj = no;
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
