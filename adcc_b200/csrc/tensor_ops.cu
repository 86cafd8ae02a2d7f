// Bandwidth-bound tensor kernels of the adccb library (sm_100a).
//
// These replace the libtensor block operations reached from libadcc_src/TensorImpl.cc:
//   strided_gather  : transpose / copy / diagonal / (anti)symmetrise   (:681-717, :667-678,
//                     :489-565, :1147-1264)
//   ew_*            : scale / add / multiply / divide / fill            (:568-664, :431-462)
//   direct_sum      : outer sum a[p] + b[q]                             (:741-789)
//   lincomb         : y = sum_k c_k x_k in ONE pass                     (:604-624)
//   dot_many        : <x, y_j> for all j in ONE pass over x             (:1120-1144)
//   spin kernels    : spin-block projector and singlet purification
//                     (libadcc_src/amplitude_vector_enforce_spin_kind.cc:33-170)
// All are HBM-bound: coalesced 16-byte accesses where alignment allows, shared-memory tiles
// for transposes, warp-shuffle + fixed-order two-stage reductions (deterministic).
#include "common.cuh"
#include <atomic>
#include <algorithm>

namespace adccb {

static inline int grid_for(long long n, int threads, int per_thread = 1) {
  long long b = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
  long long cap = 32LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------
// strided gather:  out[o] = alpha * in[off(o)] + gamma * add[o] + beta * out[o]
//   out is contiguous with shape[0..nd), off(o) = sum_d o_d * istride[d]
// ---------------------------------------------------------------------------------------
constexpr int kMaxDim = 8;
struct GatherDesc {
  int nd;
  long long shape[kMaxDim];
  long long istride[kMaxDim];
};

__global__ void gather_direct_kernel(GatherDesc d, long long n, double alpha,
                                     const double* __restrict__ in, double gamma,
                                     const double* __restrict__ add, double beta,
                                     double* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    long long rem = idx, off = 0;
#pragma unroll 1
    for (int k = d.nd - 1; k >= 0; --k) {
      const long long q = rem / d.shape[k];
      off += (rem - q * d.shape[k]) * d.istride[k];
      rem = q;
    }
    double v = alpha * in[off];
    if (add) v += gamma * add[idx];
    if (beta != 0.0) v += beta * out[idx];
    out[idx] = v;
  }
}

// tiled variant: output dim `nd-1` is the fast output axis, output dim `q` is the axis whose
// input stride is 1.  A 32x32 tile over (q, nd-1) goes through shared memory so that both the
// global read and the global write are coalesced.
__global__ void gather_tiled_kernel(GatherDesc d, int q, long long n_outer, double alpha,
                                    const double* __restrict__ in, double gamma,
                                    const double* __restrict__ add, double beta,
                                    double* __restrict__ out, long long ostride_q) {
  __shared__ double tile[32][33];
  const int last = d.nd - 1;
  const long long nl = d.shape[last], nq = d.shape[q];
  const long long tiles_l = (nl + 31) / 32, tiles_q = (nq + 31) / 32;
  const long long tiles = tiles_l * tiles_q * n_outer;
  for (long long tix = blockIdx.x; tix < tiles; tix += gridDim.x) {
    const long long tl = tix % tiles_l;
    const long long tq = (tix / tiles_l) % tiles_q;
    long long outer = tix / (tiles_l * tiles_q);
    // decode the remaining dims
    long long ioff = 0, ooff = 0, ostr = 1;
    for (int k = d.nd - 1; k >= 0; --k) {
      if (k != last && k != q) {
        const long long c = outer % d.shape[k];
        outer /= d.shape[k];
        ioff += c * d.istride[k];
        ooff += c * ostr;
      }
      ostr *= d.shape[k];
    }
    const long long l0 = tl * 32, q0 = tq * 32;
    // read: threadIdx.x runs along q (input stride 1)
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const long long l = l0 + r, qq = q0 + threadIdx.x;
      if (l < nl && qq < nq) tile[r][threadIdx.x] = in[ioff + l * d.istride[last] + qq];
    }
    __syncthreads();
    // write: threadIdx.x runs along the last output dim
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const long long qq = q0 + r, l = l0 + threadIdx.x;
      if (l < nl && qq < nq) {
        const long long o = ooff + qq * ostride_q + l;
        double v = alpha * tile[threadIdx.x][r];
        if (add) v += gamma * add[o];
        if (beta != 0.0) v += beta * out[o];
        out[o] = v;
      }
    }
    __syncthreads();
  }
}

int strided_gather(cudaStream_t st, int nd, const long long* shape, const long long* istride,
                   double alpha, const double* in, double gamma, const double* add, double beta,
                   double* out) {
  if (nd < 0 || nd > kMaxDim) return fail("strided_gather: unsupported rank");
  GatherDesc d;
  // drop extent-1 dims and merge dims that are contiguous in the input as well
  int m = 0;
  long long n = 1;
  for (int k = 0; k < nd; ++k) {
    if (shape[k] <= 0) return 0;
    n *= shape[k];
    if (shape[k] == 1) continue;
    if (m > 0 && d.istride[m - 1] == istride[k] * shape[k]) {
      d.shape[m - 1] *= shape[k];
      d.istride[m - 1] = istride[k];
    } else {
      d.shape[m] = shape[k];
      d.istride[m] = istride[k];
      ++m;
    }
  }
  if (m == 0) {
    d.shape[0] = 1;
    d.istride[0] = 0;
    m = 1;
  }
  d.nd = m;
  int q = -1;
  if (d.istride[m - 1] != 1)
    for (int k = 0; k < m - 1; ++k)
      if (d.istride[k] == 1 && d.shape[k] >= 8) q = k;
  if (q >= 0 && d.shape[m - 1] >= 8) {
    long long ostride_q = 1;
    for (int k = m - 1; k > q; --k) ostride_q *= d.shape[k];
    long long n_outer = n / (d.shape[q] * d.shape[m - 1]);
    long long tiles = ((d.shape[m - 1] + 31) / 32) * ((d.shape[q] + 31) / 32) * n_outer;
    int blocks = (int)std::min<long long>(tiles, 64LL * sm_count());
    gather_tiled_kernel<<<blocks, dim3(32, 8), 0, st>>>(d, q, n_outer, alpha, in, gamma, add, beta, out, ostride_q);
  } else {
    gather_direct_kernel<<<grid_for(n, 256, 2), 256, 0, st>>>(d, n, alpha, in, gamma, add, beta, out);
  }
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------
// elementwise family
// ---------------------------------------------------------------------------------------
enum EwOp { EW_AXPBY = 0, EW_MUL = 1, EW_DIV = 2, EW_FILL = 3, EW_DIV_SHIFT = 4, EW_ADD_SCALAR = 5 };

// out = alpha * f(x, y) + beta * out      (f depends on op; c is the scalar parameter)
template <int OP>
__global__ void ew_kernel(long long n, double alpha, const double* __restrict__ x,
                          const double* __restrict__ y, double c, double beta,
                          double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double v;
    if (OP == EW_AXPBY) v = x[i];
    else if (OP == EW_MUL) v = x[i] * y[i];
    else if (OP == EW_DIV) v = x[i] / y[i];
    else if (OP == EW_FILL) v = c;
    else if (OP == EW_DIV_SHIFT) v = x[i] / (y[i] - c);
    else v = x[i] + c;
    v *= alpha;
    if (beta != 0.0) v += beta * out[i];
    out[i] = v;
  }
}

int elementwise(cudaStream_t st, int op, long long n, double alpha, const double* x,
                const double* y, double c, double beta, double* out) {
  if (n <= 0) return 0;
  const int b = grid_for(n, 256, 4);
  switch (op) {
    case EW_AXPBY: ew_kernel<EW_AXPBY><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    case EW_MUL: ew_kernel<EW_MUL><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    case EW_DIV: ew_kernel<EW_DIV><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    case EW_FILL: ew_kernel<EW_FILL><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    case EW_DIV_SHIFT: ew_kernel<EW_DIV_SHIFT><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    case EW_ADD_SCALAR: ew_kernel<EW_ADD_SCALAR><<<b, 256, 0, st>>>(n, alpha, x, y, c, beta, out); break;
    default: return fail("elementwise: unknown op");
  }
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

// out[p * nb + q] = sa * a[p] + sb * b[q]
__global__ void direct_sum_kernel(long long na, long long nb, double sa, const double* __restrict__ a,
                                  double sb, const double* __restrict__ b, double* __restrict__ out) {
  const long long n = na * nb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / nb, q = i - p * nb;
    out[i] = sa * a[p] + sb * b[q];
  }
}

int direct_sum(cudaStream_t st, long long na, long long nb, double sa, const double* a, double sb,
               const double* b, double* out) {
  if (na * nb <= 0) return 0;
  direct_sum_kernel<<<grid_for(na * nb, 256, 4), 256, 0, st>>>(na, nb, sa, a, sb, b, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

// S[i, col0 + c] += G[i, c]  (adds a gathered column slab into the packed sigma vector)
__global__ void add_columns_kernel(long long nrows, long long ncols, double* __restrict__ S,
                                   long long lds, long long col0, const double* __restrict__ G,
                                   long long ldg) {
  const long long n = nrows * ncols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / ncols, c = idx - i * ncols;
    S[i * lds + col0 + c] += G[i * ldg + c];
  }
}

int add_columns(cudaStream_t st, long long nrows, long long ncols, double* S, long long lds,
                long long col0, const double* G, long long ldg) {
  if (nrows * ncols <= 0) return 0;
  add_columns_kernel<<<grid_for(nrows * ncols, 256, 4), 256, 0, st>>>(nrows, ncols, S, lds, col0, G, ldg);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------
// lincomb:  out = sum_k c[k] * x[k]   (k <= kMaxVec per launch, one pass, one write)
// ---------------------------------------------------------------------------------------
constexpr int kMaxVec = 32;
struct VecList {
  int n;
  const double* ptr[kMaxVec];
  double coeff[kMaxVec];
};

__global__ void lincomb_kernel(VecList v, long long n, double beta, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double s = (beta != 0.0) ? beta * out[i] : 0.0;
    for (int k = 0; k < v.n; ++k) s += v.coeff[k] * v.ptr[k][i];
    out[i] = s;
  }
}

int lincomb(cudaStream_t st, int nvec, const double* coeff, const double* const* ptrs, long long n,
            double beta, double* out) {
  if (n <= 0) return 0;
  if (nvec == 0) return elementwise(st, EW_FILL, n, beta == 0.0 ? 1.0 : 0.0, nullptr, nullptr, 0.0, beta, out);
  for (int k0 = 0; k0 < nvec; k0 += kMaxVec) {
    VecList v;
    v.n = std::min(kMaxVec, nvec - k0);
    for (int k = 0; k < v.n; ++k) {
      v.ptr[k] = ptrs[k0 + k];
      v.coeff[k] = coeff[k0 + k];
    }
    lincomb_kernel<<<grid_for(n, 256, 2), 256, 0, st>>>(v, n, k0 == 0 ? beta : 1.0, out);
    ADCCB_LAUNCH_OK();
    count_launch();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// lincomb_multi:  out_j = sum_k C[j][k] * x[k], j < nout: ONE pass over the inputs for all outputs
// (residuals / Ritz vectors / subspace collapse of the Davidson loop, adcc/solver/davidson.py:
// 139-140, 151, 168-170: nout combinations of the same 2 n_ss vectors).  Every output is summed
// in the order and with the instructions of lincomb_kernel: bit-identical to nout separate calls.
// ---------------------------------------------------------------------------------------
constexpr int kMaxOut = 8;
struct MultiList {
  int n, m;
  const double* ptr[kMaxVec];
  double* out[kMaxOut];
  double coeff[kMaxOut][kMaxVec];
};

template <int M>
__global__ void lincomb_multi_kernel(const __grid_constant__ MultiList v, long long n, double beta) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double s[M];
#pragma unroll
    for (int j = 0; j < M; ++j) s[j] = (beta != 0.0) ? beta * v.out[j][i] : 0.0;
#pragma unroll 4
    for (int k = 0; k < v.n; ++k) {
      const double x = v.ptr[k][i];
#pragma unroll
      for (int j = 0; j < M; ++j) s[j] += v.coeff[j][k] * x;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) v.out[j][i] = s[j];
  }
}

int lincomb_multi(cudaStream_t st, int nvec, int nout, const double* coeff, const double* const* ptrs,
                  long long n, double beta, double* const* outs) {
  if (n <= 0 || nout <= 0) return 0;
  if (nvec == 0) {
    for (int j = 0; j < nout; ++j) {
      const int rc = elementwise(st, EW_FILL, n, beta == 0.0 ? 1.0 : 0.0, nullptr, nullptr, 0.0, beta, outs[j]);
      if (rc) return rc;
    }
    return 0;
  }
  for (int j0 = 0; j0 < nout; j0 += kMaxOut) {
    const int m = std::min(kMaxOut, nout - j0);
    for (int k0 = 0; k0 < nvec; k0 += kMaxVec) {
      MultiList v;
      v.n = std::min(kMaxVec, nvec - k0);
      v.m = m;
      for (int k = 0; k < v.n; ++k) v.ptr[k] = ptrs[k0 + k];
      for (int j = 0; j < m; ++j) {
        v.out[j] = outs[j0 + j];
        for (int k = 0; k < v.n; ++k) v.coeff[j][k] = coeff[(size_t)(j0 + j) * nvec + k0 + k];
      }
      const double b = k0 == 0 ? beta : 1.0;
      const int grid = grid_for(n, 256, 2);
      switch (m) {
        case 1: lincomb_multi_kernel<1><<<grid, 256, 0, st>>>(v, n, b); break;
        case 2: lincomb_multi_kernel<2><<<grid, 256, 0, st>>>(v, n, b); break;
        case 3: lincomb_multi_kernel<3><<<grid, 256, 0, st>>>(v, n, b); break;
        case 4: lincomb_multi_kernel<4><<<grid, 256, 0, st>>>(v, n, b); break;
        case 5: lincomb_multi_kernel<5><<<grid, 256, 0, st>>>(v, n, b); break;
        case 6: lincomb_multi_kernel<6><<<grid, 256, 0, st>>>(v, n, b); break;
        case 7: lincomb_multi_kernel<7><<<grid, 256, 0, st>>>(v, n, b); break;
        default: lincomb_multi_kernel<8><<<grid, 256, 0, st>>>(v, n, b); break;
      }
      ADCCB_LAUNCH_OK();
      count_launch();
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// dot_many: out[j] = scale * <x, y_j>, j < nvec (<= kMaxVec per launch).  Stage 1 writes one
// partial per block and vector, stage 2 sums the partials in block order (deterministic).
// ---------------------------------------------------------------------------------------
constexpr int kDotBlocks = 4736;  // 32 x 148 (full occupancy like the other streaming kernels)
constexpr int kDotChunk = 8;      // vectors accumulated in registers per sweep

// one sweep over x and NV vectors (pointers hoisted, no predicates, 2 elements per iteration)
template <int NV>
__device__ __forceinline__ void dot_sweep(const VecList& v, int j0, const double* __restrict__ x,
                                          long long n, double (&acc)[kDotChunk]) {
  const double* p[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) p[j] = v.ptr[j0 + j];
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {
    const double x0 = x[i], x1 = x[i + stride];
    double y0[NV], y1[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) { y0[j] = p[j][i]; y1[j] = p[j][i + stride]; }
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += x0 * y0[j] + x1 * y1[j];
  }
  if (i < n) {
    const double x0 = x[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += x0 * p[j][i];
  }
}

__global__ void __launch_bounds__(256)
dot_stage1_kernel(VecList v, const double* __restrict__ x, long long n, double* __restrict__ partial) {
  __shared__ double red[8][kDotChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = 0; j0 < v.n; j0 += kDotChunk) {
    double acc[kDotChunk];
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) acc[j] = 0.0;
    switch (min(kDotChunk, v.n - j0)) {
      case 1: dot_sweep<1>(v, j0, x, n, acc); break;
      case 2: dot_sweep<2>(v, j0, x, n, acc); break;
      case 3: dot_sweep<3>(v, j0, x, n, acc); break;
      case 4: dot_sweep<4>(v, j0, x, n, acc); break;
      case 5: dot_sweep<5>(v, j0, x, n, acc); break;
      case 6: dot_sweep<6>(v, j0, x, n, acc); break;
      case 7: dot_sweep<7>(v, j0, x, n, acc); break;
      default: dot_sweep<8>(v, j0, x, n, acc); break;
    }
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) {
      double a = acc[j];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) red[warp][j] = a;
    }
    __syncthreads();
    if (threadIdx.x < kDotChunk && j0 + threadIdx.x < v.n) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      partial[(size_t)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}

// ---- bulk-copy variant: the register file cannot keep enough loads in flight for 9 concurrent
// streams (ncu: 114 regs -> 25 % occupancy, long-scoreboard stalls, 43 % of DRAM peak), so x and the NV
// vectors are streamed into a shared-memory ring with cp.async.bulk (1-D TMA, SASS UBLKCP) whose
// completion is tracked by mbarrier transaction bytes: ~180 KB in flight per SM, no register cost.
constexpr int kBulkStages = 5;
// doubles per vector per stage: few vectors -> bigger chunks, so that ~180 KB stay in flight per SM
__host__ __device__ constexpr int bulk_chunk(int nv) { return nv <= 1 ? 2048 : (nv <= 3 ? 1024 : 512); }
constexpr int kBulkThreads = 256;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ bool bar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int NV>
__global__ void __launch_bounds__(kBulkThreads + 32, 1)
dot_bulk_kernel(VecList v, int j0, const double* __restrict__ x, long long n_chunks,
                double* __restrict__ partial) {
  constexpr int kBulkChunk = bulk_chunk(NV);
  extern __shared__ __align__(128) unsigned char dsm[];
  constexpr int kStageBytes = (NV + 1) * kBulkChunk * 8;
  const uint32_t base = smem_u32(dsm);
  const uint32_t bar_base = base + kBulkStages * kStageBytes;       // full[s], then empty[s]
  __shared__ double red[kBulkThreads / 32][NV];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kBulkStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_base + 8 * s), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_base + 8 * (kBulkStages + s)),
                   "r"(kBulkThreads / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid >= kBulkThreads) {                       // producer warp
    if (tid == kBulkThreads) {
      int it = 0;
      for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const int s = it % kBulkStages;
        const uint32_t ph = (it / kBulkStages) & 1;
        while (!bar_try_wait(bar_base + 8 * (kBulkStages + s), ph ^ 1)) {}
        const uint32_t full = bar_base + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(kStageBytes)
                     : "memory");
        const uint32_t dst = base + s * kStageBytes;
        const long long off = c * kBulkChunk;
        bulk_load(dst, x + off, kBulkChunk * 8, full);
#pragma unroll
        for (int j = 0; j < NV; ++j) bulk_load(dst + (j + 1) * kBulkChunk * 8, v.ptr[j0 + j] + off, kBulkChunk * 8, full);
      }
    }
    return;
  }
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = 0.0;
  int it = 0;
  uint32_t release_bar = 0;   // stages are released one iteration late (see dgemm.cu: the reads of
                              // every lane must have completed before the refill may start)
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
    const int s = it % kBulkStages;
    const uint32_t ph = (it / kBulkStages) & 1;
    while (!bar_try_wait(bar_base + 8 * s, ph)) {}
    if (release_bar != 0) {
      __syncwarp();
      if ((tid & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(release_bar) : "memory");
    }
    release_bar = bar_base + 8 * (kBulkStages + s);
    const double* st = reinterpret_cast<const double*>(dsm + s * kStageBytes);
#pragma unroll
    for (int e = 0; e < kBulkChunk / (2 * kBulkThreads); ++e) {
      const double2 xv = reinterpret_cast<const double2*>(st)[tid + e * kBulkThreads];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double2 yv = reinterpret_cast<const double2*>(st + (j + 1) * kBulkChunk)[tid + e * kBulkThreads];
        acc[j] += xv.x * yv.x + xv.y * yv.y;
      }
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double a = acc[j];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp][j] = a;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kBulkThreads));      // consumers only (producer warp has left)
  if (tid < NV) {
    double sum = 0.0;
    for (int w = 0; w < kBulkThreads / 32; ++w) sum += red[w][tid];
    partial[(size_t)(j0 + tid) * gridDim.x + blockIdx.x] = sum;
  }
}

template <int NV>
static int launch_dot_bulk(cudaStream_t st, const VecList& v, int j0, const double* x, long long n_chunks,
                           int blocks, double* partial) {
  constexpr int smem = kBulkStages * (NV + 1) * bulk_chunk(NV) * 8 + 2 * kBulkStages * 8;
  static std::atomic<unsigned long long> attr_devices{0};      // per (function, device), see dgemm.cu
  const int dev = current_device();
  if (dev >= 64 || !((attr_devices.load(std::memory_order_acquire) >> dev) & 1ULL)) {
    ADCCB_CUDA_OK(cudaFuncSetAttribute(dot_bulk_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev < 64) attr_devices.fetch_or(1ULL << dev, std::memory_order_release);
  }
  dot_bulk_kernel<NV><<<blocks, kBulkThreads + 32, smem, st>>>(v, j0, x, n_chunks, partial);
  ADCCB_LAUNCH_OK();
  return 0;
}

__global__ void dot_stage2_kernel(const double* __restrict__ partial, int nblocks, int nvec, double scale,
                                  double beta, double* __restrict__ out) {
  const int j = blockIdx.x;
  if (j >= nvec) return;
  // fixed order: thread t sums blocks t, t+32, ...; then lanes summed by shuffle tree
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) s += partial[(size_t)j * nblocks + b];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[j] = (beta != 0.0 ? beta * out[j] : 0.0) + scale * s;
}

// tail elements [n0, n) of the bulk path, added on top of the stage-2 result
__global__ void dot_tail_kernel(VecList v, const double* __restrict__ x, long long n0, long long n,
                                double scale, double* __restrict__ out) {
  const int j = blockIdx.x;
  double s = 0.0;
  for (long long i = n0 + threadIdx.x; i < n; i += 32) s += x[i] * v.ptr[j][i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[j] += scale * s;
}

int dot_many(cudaStream_t st, int nvec, const double* x, const double* const* ys, long long n,
             double scale, double beta, double* out_dev, double* ws, long long ws_bytes) {
  if (nvec <= 0) return 0;
  if ((long long)kMaxVec * kDotBlocks * 8 > ws_bytes) return fail("dot_many: workspace too small");
  for (int k0 = 0; k0 < nvec; k0 += kMaxVec) {
    VecList v;
    v.n = std::min(kMaxVec, nvec - k0);
    bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    for (int k = 0; k < v.n; ++k) {
      v.ptr[k] = ys[k0 + k];
      aligned = aligned && (reinterpret_cast<uintptr_t>(v.ptr[k]) & 15) == 0;
    }
    if (aligned && n / 2048 >= 4LL * sm_count()) {
      // streamed through shared memory: one CTA per SM, sweeps of up to 8 vectors
      const int blocks = sm_count();
      long long done_min = n;          // elements covered by every sweep's whole chunks
      for (int j0 = 0; j0 < v.n; j0 += kDotChunk) {
        const int nv_sweep = std::min(kDotChunk, v.n - j0);
        const long long ch = bulk_chunk(nv_sweep), n_chunks = n / ch;
        int rc = 0;
        switch (nv_sweep) {
          case 1: rc = launch_dot_bulk<1>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 2: rc = launch_dot_bulk<2>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 3: rc = launch_dot_bulk<3>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 4: rc = launch_dot_bulk<4>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 5: rc = launch_dot_bulk<5>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 6: rc = launch_dot_bulk<6>(st, v, j0, x, n_chunks, blocks, ws); break;
          case 7: rc = launch_dot_bulk<7>(st, v, j0, x, n_chunks, blocks, ws); break;
          default: rc = launch_dot_bulk<8>(st, v, j0, x, n_chunks, blocks, ws); break;
        }
        if (rc) return rc;
        count_launch();
        done_min = std::min(done_min, n_chunks * ch);
        // the elements between this sweep's coverage and the common tail start are added below
        (void)done_min;
      }
      dot_stage2_kernel<<<v.n, 32, 0, st>>>(ws, blocks, v.n, scale, beta, out_dev + k0);
      ADCCB_LAUNCH_OK();
      count_launch();
      // tails: sweep with chunk ch covered [0, (n/ch)*ch); every vector's remainder is summed here
      for (int j0 = 0; j0 < v.n; j0 += kDotChunk) {
        const int nv_sweep = std::min(kDotChunk, v.n - j0);
        const long long ch = bulk_chunk(nv_sweep), n0 = (n / ch) * ch;
        if (n0 < n) {
          VecList t;
          t.n = nv_sweep;
          for (int k = 0; k < nv_sweep; ++k) t.ptr[k] = v.ptr[j0 + k];
          dot_tail_kernel<<<nv_sweep, 32, 0, st>>>(t, x, n0, n, scale, out_dev + k0 + j0);
          ADCCB_LAUNCH_OK();
          count_launch();
        }
      }
      continue;
    }
    int blocks = (int)std::min<long long>(kDotBlocks, std::max<long long>(1, (n + 255) / 256));
    dot_stage1_kernel<<<blocks, 256, 0, st>>>(v, x, n, ws);
    ADCCB_LAUNCH_OK();
    dot_stage2_kernel<<<v.n, 32, 0, st>>>(ws, blocks, v.n, scale, beta, out_dev + k0);
    ADCCB_LAUNCH_OK();
    count_launch(2);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// spin kernels on dense spin-orbital tensors (each axis = [alpha | beta])
// ---------------------------------------------------------------------------------------
struct SpinDesc {
  int nd;
  int shape[4];
  int nalpha[4];
};

// out = (T + fac * flip(T)) / 2 with spin-changing blocks zeroed.  flip relabels alpha<->beta on
// every axis.  particle axes: first nd/2 are hole (occupied) axes, the rest particle axes.
__global__ void spin_project_kernel(SpinDesc d, long long n, double fac, const double* __restrict__ in,
                                    double* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    long long rem = idx, fidx = 0, mul = 1;
    int nbeta_h = 0, nbeta_p = 0;
    for (int k = d.nd - 1; k >= 0; --k) {
      const int c = (int)(rem % d.shape[k]);
      rem /= d.shape[k];
      const bool beta = c >= d.nalpha[k];
      if (k < d.nd / 2) nbeta_h += beta; else nbeta_p += beta;
      const int fc = beta ? c - d.nalpha[k] : c + d.nalpha[k];
      fidx += fc * mul;
      mul *= d.shape[k];
    }
    out[idx] = (nbeta_h == nbeta_p) ? 0.5 * (in[idx] + fac * in[fidx]) : 0.0;
  }
}

// singlet purification: T[ia ja aa ba] = T[ia jb aa bb] + T[ia jb ab ba]; bbbb block = aaaa block
__global__ void spin_singlet_kernel(SpinDesc d, double* __restrict__ T) {
  const long long ni = d.nalpha[0], nj = d.nalpha[1], na = d.nalpha[2], nb = d.nalpha[3];
  const long long n = ni * nj * na * nb;
  const long long s2 = d.shape[3], s1 = s2 * d.shape[2], s0 = s1 * d.shape[1];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx % nb, a = (idx / nb) % na, j = (idx / (nb * na)) % nj, i = idx / (nb * na * nj);
    const double v = T[i * s0 + (j + nj) * s1 + a * s2 + (b + nb)] + T[i * s0 + (j + nj) * s1 + (a + na) * s2 + b];
    T[i * s0 + j * s1 + a * s2 + b] = v;
    T[(i + ni) * s0 + (j + nj) * s1 + (a + na) * s2 + (b + nb)] = v;
  }
}

// Fused Davidson correction step on the dense store (SURVEY K5 + K2 + K8): for a residual block x,
//   p = x / (D - shift)                                  Jacobi preconditioner (preconditioner.py:65-82)
//   q = spin projection of p (fac = +1 singlet, -1 triplet, 0 none: q = p)
//   r = index symmetrisation of q: sym_mode 0 none, 1 doubles A_ij A_ab S_(ij)(ab) (AdcMatrix.py:449-455),
//       2 CVS doubles A_ab (AdcMatrix.py:445-447)
//   out = r, and for singlets (purify) the aaaa / bbbb blocks = r[abab] + r[abba]
//         (amplitude_vector_enforce_spin_kind.cc:119-167)
// evaluated per output element from gathered inputs: one launch, one write pass, instead of
// div / project / antisymmetrise x2 / symmetrise / purify passes.
struct PrecondDesc {
  int nd, sym_mode, purify;
  int shape[4], nalpha[4];
  double shift, fac;
};
__device__ __forceinline__ double precond_q(const PrecondDesc& d, const double* __restrict__ x,
                                            const double* __restrict__ D, const int (&c)[4]) {
  long long idx = 0, fidx = 0;
  int nbeta_h = 0, nbeta_p = 0;
  for (int k = 0; k < d.nd; ++k) {
    const bool beta = c[k] >= d.nalpha[k];
    if (k < d.nd / 2) nbeta_h += beta; else nbeta_p += beta;
    idx = idx * d.shape[k] + c[k];
    fidx = fidx * d.shape[k] + (beta ? c[k] - d.nalpha[k] : c[k] + d.nalpha[k]);
  }
  const double p = D ? x[idx] / (D[idx] - d.shift) : x[idx];
  if (d.fac == 0.0) return p;
  if (nbeta_h != nbeta_p) return 0.0;
  const double pf = D ? x[fidx] / (D[fidx] - d.shift) : x[fidx];
  return 0.5 * (p + d.fac * pf);
}
__device__ __forceinline__ double precond_r(const PrecondDesc& d, const double* __restrict__ x,
                                            const double* __restrict__ D, int i, int j, int a, int b) {
  const int c0[4] = {i, j, a, b};
  if (d.nd != 4 || d.sym_mode == 0) return precond_q(d, x, D, c0);
  const int c1[4] = {i, j, b, a};
  if (d.sym_mode == 2) return 0.5 * (precond_q(d, x, D, c0) - precond_q(d, x, D, c1));
  const int c2[4] = {j, i, a, b}, c3[4] = {j, i, b, a};
  return 0.25 * (precond_q(d, x, D, c0) - precond_q(d, x, D, c2) - precond_q(d, x, D, c1) + precond_q(d, x, D, c3));
}
__global__ void precond_fused_kernel(PrecondDesc d, long long n, const double* __restrict__ x,
                                     const double* __restrict__ D, double* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    int c[4] = {0, 0, 0, 0};
    long long rem = idx;
    for (int k = d.nd - 1; k >= 0; --k) {
      c[k] = (int)(rem % d.shape[k]);
      rem /= d.shape[k];
    }
    if (d.nd == 2) {
      out[idx] = precond_q(d, x, D, c);
      continue;
    }
    double v;
    const bool b0 = c[0] >= d.nalpha[0], b1 = c[1] >= d.nalpha[1], b2 = c[2] >= d.nalpha[2], b3 = c[3] >= d.nalpha[3];
    if (d.purify && b0 == b1 && b1 == b2 && b2 == b3) {
      const int i = b0 ? c[0] - d.nalpha[0] : c[0], j = b0 ? c[1] - d.nalpha[1] : c[1];
      const int a = b0 ? c[2] - d.nalpha[2] : c[2], b = b0 ? c[3] - d.nalpha[3] : c[3];
      v = precond_r(d, x, D, i, j + d.nalpha[1], a, b + d.nalpha[3]) +
          precond_r(d, x, D, i, j + d.nalpha[1], a + d.nalpha[2], b);
    } else {
      v = precond_r(d, x, D, c[0], c[1], c[2], c[3]);
    }
    out[idx] = v;
  }
}

// Spin expansion of a restricted (spatial) two-electron block in chemists' order:
// out[p,q,r,s] = T[p',q',r',s'] if spin(p)==spin(q) and spin(r)==spin(s), else 0, where every output
// axis is [alpha | beta] over the same n[k] spatial orbitals (adcc/backends/EriBuilder.py:117-119).
__global__ void spin_expand4_kernel(int n0, int n1, int n2, int n3, const double* __restrict__ T,
                                    double* __restrict__ out) {
  const long long total = 16LL * n0 * n1 * n2 * n3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long rem = idx;
    const int s = (int)(rem % (2 * n3)); rem /= 2 * n3;
    const int r = (int)(rem % (2 * n2)); rem /= 2 * n2;
    const int q = (int)(rem % (2 * n1)); rem /= 2 * n1;
    const int p = (int)rem;
    const bool bp = p >= n0, bq = q >= n1, br = r >= n2, bs = s >= n3;
    double v = 0.0;
    if (bp == bq && br == bs)
      v = T[(((long long)(bp ? p - n0 : p) * n1 + (bq ? q - n1 : q)) * n2 + (br ? r - n2 : r)) * n3 +
            (bs ? s - n3 : s)];
    out[idx] = v;
  }
}

int spin_expand4(cudaStream_t st, const int* n, const double* T, double* out) {
  const long long total = 16LL * n[0] * n[1] * n[2] * n[3];
  if (total <= 0) return 0;
  spin_expand4_kernel<<<grid_for(total, 256, 2), 256, 0, st>>>(n[0], n[1], n[2], n[3], T, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int spin_project(cudaStream_t st, int nd, const int* shape, const int* nalpha, double fac,
                 const double* in, double* out) {
  if (nd != 2 && nd != 4) return fail("spin_project: rank must be 2 or 4");
  if (in == out) return fail("spin_project: in-place not supported");
  SpinDesc d;
  d.nd = nd;
  long long n = 1;
  for (int k = 0; k < nd; ++k) {
    d.shape[k] = shape[k];
    d.nalpha[k] = nalpha[k];
    if (2 * nalpha[k] != shape[k]) return fail("spin_project: needs equal alpha/beta halves");
    n *= shape[k];
  }
  if (n == 0) return 0;
  spin_project_kernel<<<grid_for(n, 256, 2), 256, 0, st>>>(d, n, fac, in, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int precond_apply(cudaStream_t st, int nd, const int* shape, const int* nalpha, const double* x, const double* D,
                  double shift, int sym_mode, double spin_fac, int purify, double* out) {
  if (nd != 2 && nd != 4) return fail("precond_apply: rank 2 or 4 blocks");
  if (sym_mode < 0 || sym_mode > 2) return fail("precond_apply: unknown symmetrisation mode");
  PrecondDesc d;
  d.nd = nd; d.sym_mode = (nd == 4) ? sym_mode : 0; d.purify = (nd == 4) ? purify : 0;
  d.shift = shift; d.fac = spin_fac;
  long long n = 1;
  for (int k = 0; k < 4; ++k) {
    d.shape[k] = k < nd ? shape[k] : 1;
    d.nalpha[k] = k < nd ? nalpha[k] : 1;
    if (k < nd) n *= shape[k];
    if (k < nd && (spin_fac != 0.0 || purify) && 2 * nalpha[k] != shape[k])
      return fail("precond_apply: spin kinds need equal alpha/beta halves");
  }
  if (nd == 4 && sym_mode == 1 && (shape[0] != shape[1] || shape[2] != shape[3]))
    return fail("precond_apply: antisymmetrised index pairs need equal extents");
  if (nd == 4 && sym_mode == 2 && shape[2] != shape[3])
    return fail("precond_apply: antisymmetrised index pairs need equal extents");
  if (n == 0) return 0;
  precond_fused_kernel<<<grid_for(n, 256, 2), 256, 0, st>>>(d, n, x, D, out);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int spin_singlet(cudaStream_t st, const int* shape, const int* nalpha, double* T) {
  SpinDesc d;
  d.nd = 4;
  long long n = 1;
  for (int k = 0; k < 4; ++k) {
    d.shape[k] = shape[k];
    d.nalpha[k] = nalpha[k];
    if (2 * nalpha[k] != shape[k]) return fail("spin_singlet: needs equal alpha/beta halves");
    n *= nalpha[k];
  }
  if (n == 0) return 0;
  spin_singlet_kernel<<<grid_for(n, 256, 1), 256, 0, st>>>(d, T);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace adccb
