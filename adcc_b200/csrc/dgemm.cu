// FP64 contraction core:  C[m,n] = alpha * sum_k A[m,k] * B[n,k] + beta * C[m,n]
//
// Replaces libtensor's block-wise DGEMM evaluation reached from
// libadcc_src/TensorImpl.cc:856-876,901-1117 (tensordot -> lt::contract).
//
// Fast path (dgemm_tn_tma_kernel): both operands K-contiguous, 16-byte aligned rows.
//   * one producer warp streams 16-wide K slabs of A and B into a STAGES-deep shared-memory
//     ring with TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, zero fill out of bounds),
//     completion tracked by mbarrier transaction counts;
//   * consumer warps read DMMA fragments with conflict-free LDS.128 (row permutation inside
//     each 8-row block + the 128B swizzle) and issue DMMA.8x8x4 (the only FP64 tensor shape
//     sm_100a has natively; every larger mma.sync f64 shape lowers to it);
//   * accumulators stay in registers; the epilogue applies alpha/beta or writes split-K
//     partials that a second kernel reduces in a fixed order (deterministic).
// N-major B (template flag BNM): the same kernel with B given as B[k][n] (n contiguous), i.e.
//   C = A . B without a transposed copy of B.  The B slab of a stage is then BN/16 TMA boxes of
//   16 k-rows x 16 n-columns; the k <-> (h, t, slot) relabelling of the A fragments puts the four
//   k-rows of one DMMA on rows 2t (+1) of an 8-row swizzle period, so the 64-bit fragment loads
//   of a half warp fall into 8 distinct 16-byte chunks (conflict-free).  Used for the ovvv
//   couplings: u1 . V2 and Ue . V2^T both read the ONE stored layout V2[a,(j,BC)].
// Fallback path (dgemm_generic_kernel): arbitrary element strides, used when the TMA
// alignment rules do not hold (odd leading dimension) - same DMMA core, plain loads.
#include "common.cuh"
#include <vector>
#include <mutex>
#include <atomic>

namespace adccb {

// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------
// TMA + DMMA kernel
// ---------------------------------------------------------------------------------------
constexpr int kBK = 16;  // doubles per K slab = one 128-byte swizzle row

template <int WARPS_M, int WARPS_N, int MB, int NB, int STAGES>
struct TileCfg {
  static constexpr int kConsumerWarps = WARPS_M * WARPS_N;
  static constexpr int kThreads = (kConsumerWarps + 1) * 32;
  static constexpr int BM = WARPS_M * MB * 8;
  static constexpr int BN = WARPS_N * NB * 8;
  static constexpr int A_BYTES = BM * kBK * 8;
  static constexpr int B_BYTES = BN * kBK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // a TMA box holds at most 256 rows: wider B tiles arrive as two boxes (each a multiple of the
  // 8-row swizzle period, so the second one lands 1024-byte aligned behind the first)
  static constexpr int B_BOXES = BN <= 256 ? 1 : 2;
  static constexpr int B_BOX_ROWS = BN / B_BOXES;
  static_assert(BN % B_BOXES == 0 && B_BOX_ROWS % 8 == 0 && B_BOX_ROWS <= 256, "B tile does not split into TMA boxes");
  static_assert(A_BYTES % 1024 == 0 && STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
};

// Additional destinations of every finished C tile: the same (row, col) of buffers in PEER GPUs'
// memory (mapped through CUDA IPC, written over NVLink).  This fuses the exchange step that follows
// a column-slab GEMM (all-gather of the slabs) into the epilogue: the tile goes out while the
// other CTAs are still in their DMMA loops, so no collective is left on the critical path.
constexpr int kMaxPeers = 7;
struct PeerDests {
  int n;
  double* ptr[kMaxPeers];
};

// epilogue shared by the main kernel and the stream-K fix-up: thread (warp, lane) holds
// C[row pg][cols t, t+4] of each 8x8 block of its warp tile
// NAT: the B fragments were read with lane g <-> tile column g (N-major B), so the thread holds
// columns 2t, 2t+1 of each 8x8 block instead of t, t+4
template <int WARPS_N, int MB, int NB, int BM, int BN, bool PEERS, bool NAT = false>
__device__ __forceinline__ void store_tile(const double (&acc)[MB][NB][2], int warp, int lane,
                                           int tile_m, int tile_n, double* __restrict__ C,
                                           long long ldc, int M, int N, double alpha, double beta,
                                           const PeerDests& peers) {
  if constexpr (PEERS) {   // beta == 0 by contract; one pass per destination keeps the stores streaming
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) | ((g & 1) << 2);
    const int row0 = tile_m * BM + wm * MB * 8 + pg;
    const int col0 = tile_n * BN + wn * NB * 8 + t;
    for (int d = -1; d < peers.n; ++d) {
      double* dst = d < 0 ? C : peers.ptr[d];
#pragma unroll
      for (int i = 0; i < MB; ++i) {
        const int r = row0 + i * 8;
        if (r >= M) continue;
        double* Crow = dst + (size_t)r * ldc;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int c = col0 + j * 8;
          if (c < N) Crow[c] = alpha * acc[i][j][0];
          if (c + 4 < N) Crow[c + 4] = alpha * acc[i][j][1];
        }
      }
    }
    return;
  }
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const int pg = (g >> 1) | ((g & 1) << 2);
  const int row0 = tile_m * BM + wm * MB * 8 + pg;
  const int col0 = tile_n * BN + wn * NB * 8 + (NAT ? 2 * t : t);
  constexpr int kSecond = NAT ? 1 : 4;
#pragma unroll
  for (int i = 0; i < MB; ++i) {
    const int r = row0 + i * 8;
    if (r >= M) continue;
    double* Crow = C + (size_t)r * ldc;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int c = col0 + j * 8;
      if (c < N) Crow[c] = (beta == 0.0) ? alpha * acc[i][j][0] : alpha * acc[i][j][0] + beta * Crow[c];
      if (c + kSecond < N)
        Crow[c + kSecond] =
            (beta == 0.0) ? alpha * acc[i][j][1] : alpha * acc[i][j][1] + beta * Crow[c + kSecond];
    }
  }
}

// Work distribution (persistent, one CTA per SM): tiles are ordered M-fastest.
//   * data-parallel phase: `dp_waves` rounds in which CTA b owns the whole tile  w * gridDim.x + b
//     (the CTAs running together cover neighbouring tiles, so A/B slabs are shared through L2);
//   * stream-K phase: the remaining tiles' (tile, k-slab) space is cut into equal contiguous
//     spans, one per CTA.  A span segment that covers a tile completely stores it directly; the
//     (at most two) partially covered tiles at the ends of a span go to `partial` (slot 2*cta for
//     the span's first segment, 2*cta+1 for its last) and dgemm_fixup_kernel sums the segments of
//     each such tile in CTA order (deterministic).
// One scheme thus covers plain tiling, split-K for skinny outputs and the partial last wave.
struct SegIter {
  int w, dp_waves, tiles_m, k_iters, rem, k_next;
  long long tile_next;
  bool first_sk;
  __device__ __forceinline__ void init(int dp_waves_, int tiles_m_, int k_iters_, long long sk_tile0,
                                       long long sk_total, long long ipc) {
    w = 0; dp_waves = dp_waves_; tiles_m = tiles_m_; k_iters = k_iters_;
    const long long g0 = (long long)blockIdx.x * ipc;
    const long long g1 = min(g0 + ipc, sk_total);
    rem = g1 > g0 ? (int)(g1 - g0) : 0;
    const long long tl = g0 / k_iters;
    tile_next = sk_tile0 + tl;
    k_next = (int)(g0 - tl * k_iters);
    first_sk = true;
  }
  // returns false when the CTA has no more work
  __device__ __forceinline__ bool next(int& tm, int& tn, int& k_start, int& n_seg, int& slot) {
    long long tile;
    if (w < dp_waves) {
      tile = (long long)w * gridDim.x + blockIdx.x;
      ++w;
      k_start = 0; n_seg = k_iters; slot = -1;
    } else if (rem > 0) {
      tile = tile_next++;
      k_start = k_next;
      n_seg = min(rem, k_iters - k_next);
      slot = (n_seg == k_iters) ? -1 : (first_sk ? 0 : 1);
      rem -= n_seg; k_next = 0; first_sk = false;
    } else {
      return false;
    }
    tm = (int)(tile % tiles_m);
    tn = (int)(tile / tiles_m);
    return true;
  }
};

template <int WARPS_M, int WARPS_N, int MB, int NB, int STAGES, bool PEERS, bool BNM = false>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + 1) * 32, 1)
dgemm_tn_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB, double* __restrict__ C,
                    long long ldc, int M, int N, int tiles_m, int k_iters, int dp_waves,
                    long long sk_tile0, long long sk_total, long long ipc, double alpha, double beta,
                    double* __restrict__ partial, const __grid_constant__ PeerDests peers) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, MB, NB, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // full[s] at bar_base + 8*s, empty[s] at bar_base + 8*(STAGES+s)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), Cfg::kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  SegIter seg;
  seg.init(dp_waves, tiles_m, k_iters, sk_tile0, sk_total, ipc);
  int tm, tn, k_start, n_seg, slot;
  int it = 0;

  if (warp == Cfg::kConsumerWarps) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      while (seg.next(tm, tn, k_start, n_seg, slot)) {
        for (int n = 0; n < n_seg; ++n, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_base + 8 * (STAGES + s), ph ^ 1);
          const uint32_t full = bar_base + 8 * s;
          mbar_expect_tx(full, Cfg::STAGE_BYTES);
          const uint32_t dst = smem_base + s * Cfg::STAGE_BYTES;
          tma_load_2d(dst, &tmA, full, (k_start + n) * kBK, tm * Cfg::BM);
          if constexpr (BNM) {
            static_assert(!BNM || Cfg::BN % 16 == 0, "N-major B needs whole 16-column boxes");
            for (int b = 0; b < Cfg::BN / 16; ++b)
              tma_load_2d(dst + Cfg::A_BYTES + b * 2048, &tmB, full, tn * Cfg::BN + 16 * b, (k_start + n) * kBK);
          } else {
#pragma unroll
            for (int b = 0; b < Cfg::B_BOXES; ++b)
              tma_load_2d(dst + Cfg::A_BYTES + b * Cfg::B_BOX_ROWS * 128, &tmB, full, (k_start + n) * kBK,
                          tn * Cfg::BN + b * Cfg::B_BOX_ROWS);
          }
        }
      }
    }
    return;
  }

  // ===== DMMA consumers =====
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  // fragment row g of every 8-row block lives in tile row pg: lanes of one quarter warp then
  // hit rows r and r^4, whose swizzled 16B chunks fall into disjoint bank groups
  const int pg = (g >> 1) | ((g & 1) << 2);
  const uint32_t a_off = (uint32_t)((wm * MB * 8 + pg) * 128 + ((t ^ pg) << 4));
  const uint32_t b_off = (uint32_t)(Cfg::A_BYTES + (wn * NB * 8 + pg) * 128 + ((t ^ pg) << 4));
  // N-major B: column (wn*NB*8 + 8j + g) of the tile = box col>>4, 16-byte chunk (col&15)>>1, half col&1;
  // k-row 8h + 2t + q of the slab sits at chunk ^ (2t + q)
  uint32_t bn_off[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int col = wn * NB * 8 + 8 * j + g;
    bn_off[j] = (uint32_t)(Cfg::A_BYTES + (col >> 4) * 2048 + (2 * t) * 128 + ((col & 1) << 3));
  }

  double acc[MB][NB][2];
  // A stage is handed back to the producer ONE iteration late, after the wait for the next
  // stage: by then every DMMA that consumed the stage's fragments has been issued, so all of its
  // shared-memory reads (of every lane) have completed.  Releasing right behind the last LDS is
  // not enough: the arrive of lane 0 does not order the other lanes' reads, and the TMA refill of
  // an L2-resident slab was observed to overtake a delayed read (wrong rows of the last-loaded
  // fragment, about once per 10^4 tiles, first tiles of a launch).
  uint32_t release_bar = 0;
  while (seg.next(tm, tn, k_start, n_seg, slot)) {
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int n = 0; n < n_seg; ++n, ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(bar_base + 8 * s, ph);
      if (release_bar != 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(release_bar);
      }
      release_bar = bar_base + 8 * (STAGES + s);
      const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES + a_off;
      const uint32_t sb = smem_base + s * Cfg::STAGE_BYTES + b_off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // logical 16B chunk t + 4h of the row: doubles k = 8h + 2t, 8h + 2t + 1.  The same
        // k <-> (h, t, slot) relabelling is used for A and B, which leaves the dot product intact.
        double2 af[MB], bf[NB];
#pragma unroll
        for (int i = 0; i < MB; ++i) af[i] = lds128((sa + i * 1024) ^ (h << 6));
        if constexpr (BNM) {
          const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES + h * 1024;   // rows 8h ..
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            // (col & 15) >> 1 = 4 * (j & 1) + (g >> 1) when the warp's first column is a multiple of 16,
            // else shifted by 4: computed from the column itself
            const int col = wn * NB * 8 + 8 * j + g;
            const uint32_t c = (uint32_t)((col & 15) >> 1);
            const uint32_t base = stage + bn_off[j];
            bf[j].x = lds64(base + ((c ^ (uint32_t)(2 * t)) << 4));
            bf[j].y = lds64(base + 128 + ((c ^ (uint32_t)(2 * t + 1)) << 4));
          }
        } else {
#pragma unroll
          for (int j = 0; j < NB; ++j) bf[j] = lds128((sb + j * 1024) ^ (h << 6));
        }
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
      }
    }

    if (slot < 0) {
      store_tile<WARPS_N, MB, NB, Cfg::BM, Cfg::BN, PEERS, BNM>(acc, warp, lane, tm, tn, C, ldc, M, N, alpha, beta, peers);
    } else {
      // fragment-ordered partial: element (reg, consumer thread) -> coalesced
      constexpr int kCT = Cfg::kConsumerWarps * 32;
      double* P = partial + ((size_t)2 * blockIdx.x + slot) * (size_t)(MB * NB * 2) * kCT;
      const int ctid = threadIdx.x;
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          P[(size_t)((i * NB + j) * 2 + 0) * kCT + ctid] = acc[i][j][0];
          P[(size_t)((i * NB + j) * 2 + 1) * kCT + ctid] = acc[i][j][1];
        }
    }
  }
}

// sums the partial segments of every stream-K tile that no single CTA covered, in CTA order.
// blockIdx.y picks ONE 8x8 fragment block (i, j) of the tile: a split-K GEMM with few output tiles
// (skinny ovvv / ooov streams: 4 tiles cut over 148 CTAs) then still fills the machine instead of
// summing 37 segments x 20 registers in 4 CTAs (measured 0.10 ms of a 4 ms GEMM).
template <int WARPS_M, int WARPS_N, int MB, int NB, int STAGES, bool PEERS, bool BNM = false>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
dgemm_fixup_kernel(double* __restrict__ C, long long ldc, int M, int N, int tiles_m, int k_iters,
                   long long sk_tile0, long long sk_tiles, long long ipc, double alpha, double beta,
                   const double* __restrict__ partial, const __grid_constant__ PeerDests peers) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, MB, NB, STAGES>;
  constexpr int kThreads = Cfg::kConsumerWarps * 32;
  const int bi = blockIdx.y / NB, bj = blockIdx.y % NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const int pg = (g >> 1) | ((g & 1) << 2);
  for (long long tl = blockIdx.x; tl < sk_tiles; tl += gridDim.x) {
    const long long t0 = tl * (long long)k_iters, t1 = t0 + k_iters;
    const long long b_first = t0 / ipc, b_last = (t1 - 1) / ipc;
    if (b_first == b_last && b_first * ipc <= t0 && (b_first + 1) * ipc >= t1) continue;  // stored directly
    // CTA b's span starts inside this tile (its first segment, slot 0) or before it (then the tile holds
    // its last segment, slot 1).  Four independent loads per step: a skinny split-K GEMM sums up to
    // one segment per SM here, and a dependent chain of 148 x 2 loads took 0.09 ms
    // (profiles/r02_bench_launches.csv); the order of the additions stays fixed (deterministic).
    double a0 = 0.0, a1 = 0.0;
    const size_t seg = (size_t)(MB * NB * 2) * kThreads;
    const double* P0 = partial + (size_t)((bi * NB + bj) * 2) * kThreads + threadIdx.x;
    long long b = b_first;
    for (; b + 3 <= b_last; b += 4) {
      double x[4], y[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int slot = ((b + q) * ipc >= t0) ? 0 : 1;
        const double* P = P0 + ((size_t)2 * (b + q) + slot) * seg;
        x[q] = P[0];
        y[q] = P[kThreads];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { a0 += x[q]; a1 += y[q]; }
    }
    for (; b <= b_last; ++b) {
      const int slot = (b * ipc >= t0) ? 0 : 1;
      const double* P = P0 + ((size_t)2 * b + slot) * seg;
      a0 += P[0];
      a1 += P[kThreads];
    }
    const long long tile = sk_tile0 + tl;
    const int r = (int)(tile % tiles_m) * Cfg::BM + wm * MB * 8 + pg + bi * 8;
    const int c = (int)(tile / tiles_m) * Cfg::BN + wn * NB * 8 + (BNM ? 2 * t : t) + bj * 8;
    constexpr int kSecond = BNM ? 1 : 4;
    if (r >= M) continue;
    for (int d = -1; d < (PEERS ? peers.n : 0); ++d) {
      double* Crow = (d < 0 ? C : peers.ptr[d]) + (size_t)r * ldc;
      if (c < N) Crow[c] = (beta == 0.0) ? alpha * a0 : alpha * a0 + beta * Crow[c];
      if (c + kSecond < N) Crow[c + kSecond] = (beta == 0.0) ? alpha * a1 : alpha * a1 + beta * Crow[c + kSecond];
    }
  }
}

// ---------------------------------------------------------------------------------------
// generic fallback: arbitrary element strides (A(m,k) = A[m*ars + k*acs], same for B)
// 64x64 tile, 4 warps (2x2), each warp 32x32, BK = 16, plain loads -> padded smem -> DMMA
// ---------------------------------------------------------------------------------------
constexpr int GB = 64;        // tile edge
constexpr int GPITCH = 18;    // doubles per smem row (16 + 2 pad; keeps 16B alignment)

__global__ void __launch_bounds__(128)
dgemm_generic_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                     long long ars, long long acs, const double* __restrict__ B, long long brs,
                     long long bcs, double beta, double* __restrict__ C, long long ldc) {
  __shared__ __align__(16) double sA[GB * GPITCH];
  __shared__ __align__(16) double sB[GB * GPITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * GB, n0 = blockIdx.y * GB;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int k0 = 0; k0 < K; k0 += kBK) {
    for (int e = threadIdx.x; e < GB * kBK; e += 128) {
      int r, k;
      // pick the traversal order that makes the contiguous operand dimension fastest
      if (acs == 1) { r = e / kBK; k = e % kBK; } else { k = e / GB; r = e % GB; }
      double v = 0.0;
      if (m0 + r < M && k0 + k < K) v = A[(long long)(m0 + r) * ars + (long long)(k0 + k) * acs];
      sA[r * GPITCH + k] = v;
      if (bcs == 1) { r = e / kBK; k = e % kBK; } else { k = e / GB; r = e % GB; }
      v = 0.0;
      if (n0 + r < N && k0 + k < K) v = B[(long long)(n0 + r) * brs + (long long)(k0 + k) * bcs];
      sB[r * GPITCH + k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = sA[(wm * 32 + i * 8 + g) * GPITCH + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = sB[(wn * 32 + j * 8 + g) * GPITCH + kk + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm * 32 + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = n0 + wn * 32 + j * 8 + 2 * t + q;
        if (c >= N) continue;
        double* dst = C + (long long)r * ldc + c;
        *dst = (beta == 0.0) ? alpha * acc[i][j][q] : alpha * acc[i][j][q] + beta * (*dst);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static bool make_map_2d(CUtensorMap* tm, const double* base, long long rows, long long cols,
                        long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// B[k][n], n contiguous: boxes of 16 n-columns (one 128-byte swizzle row) x 16 k-rows
static bool make_map_nmajor(CUtensorMap* tm, const double* base, long long k_rows, long long n_cols, long long ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)n_cols, (cuuint64_t)k_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, (cuuint32_t)kBK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int WARPS_M, int WARPS_N, int MB, int NB, int STAGES, bool PEERS, bool BNM = false>
static int launch_tma(cudaStream_t st, int M, int N, int K, double alpha, const double* A,
                      long long lda, const double* B, long long ldb, double beta, double* C,
                      long long ldc, double* ws, long long ws_bytes, const PeerDests& peers) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, MB, NB, STAGES>;
  auto kern = dgemm_tn_tma_kernel<WARPS_M, WARPS_N, MB, NB, STAGES, PEERS, BNM>;
  // the attribute belongs to the (function, device) pair: set it once per device
  // (one-process-many-GPU hosts launch the same instantiation on several devices)
  static std::atomic<unsigned long long> attr_devices{0};
  const int dev = current_device();
  if (dev >= 64 || !((attr_devices.load(std::memory_order_acquire) >> dev) & 1ULL)) {
    ADCCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (dev < 64) attr_devices.fetch_or(1ULL << dev, std::memory_order_release);
  }
  CUtensorMap tmA, tmB;
  // BNM: ldb is the stride between k-rows of B[k][n]
  if (!make_map_2d(&tmA, A, M, K, lda, Cfg::BM) ||
      !(BNM ? make_map_nmajor(&tmB, B, K, N, ldb) : make_map_2d(&tmB, B, N, K, ldb, Cfg::B_BOX_ROWS)))
    return fail("cuTensorMapEncodeTiled failed");
  const int tm = (M + Cfg::BM - 1) / Cfg::BM, tn = (N + Cfg::BN - 1) / Cfg::BN;
  const long long tiles = (long long)tm * tn;
  const int k_iters = (K + kBK - 1) / kBK;
  const int nsm = sm_count();
  const long long slot_bytes = (long long)MB * NB * 2 * Cfg::kConsumerWarps * 32 * 8;
  const bool can_sk = ws != nullptr && 2LL * nsm * slot_bytes <= ws_bytes;
  // data-parallel rounds of whole tiles, stream-K for the rest (see SegIter)
  long long n_ctas, sk_tile0, sk_tiles, ipc;
  int dp_waves;
  if (!can_sk) {                       // whole tiles only, one CTA per tile
    n_ctas = tiles; dp_waves = 1; sk_tile0 = tiles; sk_tiles = 0; ipc = k_iters;
  } else if (tiles >= nsm) {
    const long long q = tiles / nsm, r = tiles % nsm;
    n_ctas = nsm;
    // The ragged last round alone goes to stream-K when it gives every CTA at least half a tile;
    // otherwise the last full round joins it ("two-tile" stream-K: spans of 1 - 1.5 tiles).  Stream-K
    // spans start at unrelated k offsets, so nothing they read is shared through L2: with the
    // two-tile rule everywhere the vvvv GEMM (29.5 rounds) read 100.6 GB per launch, with this rule
    // 77.7 GB, same duration (profiles/r02_gemm_pacing_and_streamk_tail_ab.txt).
    dp_waves = (int)((r == 0 || 2 * r >= nsm) ? q : q - 1);
    sk_tile0 = (long long)dp_waves * nsm;
    sk_tiles = tiles - sk_tile0;
    ipc = (sk_tiles * k_iters + n_ctas - 1) / n_ctas;
  } else {                             // fewer tiles than SMs: split K, spans of >= 8 k-slabs
    const long long total = tiles * k_iters;
    n_ctas = std::min<long long>(nsm, std::max<long long>(tiles, total / 8));
    dp_waves = 0; sk_tile0 = 0; sk_tiles = tiles;
    ipc = (total + n_ctas - 1) / n_ctas;
    n_ctas = (total + ipc - 1) / ipc;
  }
  const long long sk_total = sk_tiles * k_iters;
  const bool has_partials = sk_tiles > 0 && (ipc % k_iters) != 0;
  kern<<<(unsigned)n_ctas, Cfg::kThreads, Cfg::SMEM_BYTES, st>>>(tmA, tmB, C, ldc, M, N, tm, k_iters, dp_waves,
                                                                 sk_tile0, sk_total, ipc, alpha, beta, ws, peers);
  ADCCB_LAUNCH_OK();
  count_launch();
  if (has_partials) {
    const dim3 blocks((unsigned)std::min<long long>(sk_tiles, 4LL * nsm), MB * NB);
    dgemm_fixup_kernel<WARPS_M, WARPS_N, MB, NB, STAGES, PEERS, BNM><<<blocks, Cfg::kConsumerWarps * 32, 0, st>>>(
        C, ldc, M, N, tm, k_iters, sk_tile0, sk_tiles, ipc, alpha, beta, ws, peers);
    ADCCB_LAUNCH_OK();
    count_launch();
  }
  return 0;
}

static int dgemm_impl(cudaStream_t st, long long M, long long N, long long K, double alpha,
                      const double* A, long long ars, long long acs, const double* B, long long brs,
                      long long bcs, double beta, double* C, long long ldc, double* ws,
                      long long ws_bytes, int tile_hint, const PeerDests& peers) {
  if (M <= 0 || N <= 0) return 0;
  if (peers.n > 0 && (beta != 0.0 || K <= 0)) return fail("dgemm_peers: needs beta == 0 and K > 0");
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return fail("dgemm: extent exceeds int32");
  if (K <= 0) {
    // C = beta * C
    dim3 grid((unsigned)((M + GB - 1) / GB), (unsigned)((N + GB - 1) / GB));
    dgemm_generic_kernel<<<grid, 128, 0, st>>>((int)M, (int)N, 0, alpha, A, ars, acs, B, brs, bcs, beta, C, ldc);
    ADCCB_LAUNCH_OK();
    count_launch();
    return 0;
  }
  const bool tma_ok = acs == 1 && bcs == 1 && (ars % 2 == 0) && (brs % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && get_encode_fn() != nullptr &&
                      (peers.n > 0 || (long long)M * N * K >= 32LL * 32 * 32);
  // C = A . B with B[k][n] stored n-contiguous (no transposed copy): same tiles, N-major B loads
  const bool tma_nn_ok = peers.n == 0 && acs == 1 && brs == 1 && N > 1 && (ars % 2 == 0) && (bcs % 2 == 0) &&
                         ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && get_encode_fn() != nullptr &&
                         (long long)M * N * K >= 32LL * 32 * 32;
  if (tma_nn_ok && !(bcs == 1)) {
#define ADCCB_TMA_NN(WM, WN, MB, NB)                                                                       \
  return launch_tma<WM, WN, MB, NB, 6, false, true>(st, (int)M, (int)N, (int)K, alpha, A, ars, B, bcs, beta, C, \
                                                    ldc, ws, ws_bytes, peers)
    if (tile_hint == 16 || (tile_hint == 0 && M <= 16)) ADCCB_TMA_NN(1, 8, 2, 2);
    if (tile_hint == 40 || (tile_hint == 0 && M <= 40)) ADCCB_TMA_NN(1, 8, 5, 2);
    if (tile_hint == 64 || (tile_hint == 0 && M <= 64)) ADCCB_TMA_NN(1, 8, 8, 2);
    ADCCB_TMA_NN(2, 4, 8, 4);
#undef ADCCB_TMA_NN
  }
  if (tma_ok) {
    // row-tile height: the smallest tile that covers M for skinny outputs (these are HBM-bound
    // streams over the big operand, so DMMA rows spent on zero padding would make them
    // tensor-bound instead), else whichever of 112 / 128 pads M less
    // (e.g. 780 occupied pairs -> 7 x 112 = 784 instead of 7 x 128 = 896)
#define ADCCB_TMA(WM, WN, MB, NB)                                                                    \
  return peers.n > 0 ? launch_tma<WM, WN, MB, NB, 6, true>(st, (int)M, (int)N, (int)K, alpha, A, ars, B, \
                                                           brs, beta, C, ldc, ws, ws_bytes, peers)      \
                     : launch_tma<WM, WN, MB, NB, 6, false>(st, (int)M, (int)N, (int)K, alpha, A, ars, B, \
                                                            brs, beta, C, ldc, ws, ws_bytes, peers)
    // skinny split-K stream against a B operand of up to 400 rows (Ue . V2^T of the ovvv coupling at
    // nvirt = 400: M = nocc, K = nocc * npv): ONE 40 x 400 tile, 10 consumer warps of 5 x 5 fragment
    // blocks, instead of four 128-column tiles of which the last is 88 % padding (up to 16 rows the
    // 16-row tiles stay: these streams are HBM-bound only while few DMMA rows are spent on padding)
    if (peers.n == 0 && ((tile_hint == 0 && M > 16 && M <= 40 && N > 256 && N <= 400) || tile_hint == 400))
      return launch_tma<1, 10, 5, 5, 4, false>(st, (int)M, (int)N, (int)K, alpha, A, ars, B, brs, beta, C, ldc, ws,
                                               ws_bytes, peers);
    if (tile_hint == 0 || tile_hint == 16 || tile_hint == 40 || tile_hint == 64) {
      if (tile_hint == 16 || (tile_hint == 0 && M <= 16)) ADCCB_TMA(1, 8, 2, 2);
      if (tile_hint == 40 || (tile_hint == 0 && M <= 40)) ADCCB_TMA(1, 8, 5, 2);
      if (tile_hint == 64 || (tile_hint == 0 && M <= 64)) ADCCB_TMA(1, 8, 8, 2);
    }
    const long long pad112 = (M + 111) / 112 * 112, pad128 = (M + 127) / 128 * 128;
    if (tile_hint == 112 || (tile_hint == 0 && pad112 < pad128)) ADCCB_TMA(1, 8, 14, 2);
    ADCCB_TMA(2, 4, 8, 4);
#undef ADCCB_TMA
  }
  if (peers.n > 0) return fail("dgemm_peers: operands must be K-contiguous and 16-byte aligned (TMA path)");
  dim3 grid((unsigned)((M + GB - 1) / GB), (unsigned)((N + GB - 1) / GB));
  if (grid.y > 65535) return fail("dgemm(generic): too many N tiles");
  dgemm_generic_kernel<<<grid, 128, 0, st>>>((int)M, (int)N, (int)K, alpha, A, ars, acs, B, brs, bcs, beta, C, ldc);
  ADCCB_LAUNCH_OK();
  count_launch();
  return 0;
}

int dgemm_strided(cudaStream_t st, long long M, long long N, long long K, double alpha,
                  const double* A, long long ars, long long acs, const double* B, long long brs,
                  long long bcs, double beta, double* C, long long ldc, double* ws,
                  long long ws_bytes, int tile_hint) {
  PeerDests none;
  none.n = 0;
  return dgemm_impl(st, M, N, K, alpha, A, ars, acs, B, brs, bcs, beta, C, ldc, ws, ws_bytes,
                    tile_hint, none);
}

// C = alpha * A B^T written to the local C AND to the same offsets of n_peers peer buffers
int dgemm_peers(cudaStream_t st, long long M, long long N, long long K, double alpha,
                const double* A, long long lda, const double* B, long long ldb, double* C,
                int n_peers, double* const* peer_C, long long ldc, double* ws, long long ws_bytes,
                int tile_hint) {
  if (n_peers < 0 || n_peers > kMaxPeers) return fail("dgemm_peers: at most 7 peer destinations");
  PeerDests peers;
  peers.n = n_peers;
  for (int d = 0; d < n_peers; ++d) {
    if (peer_C[d] == nullptr) return fail("dgemm_peers: null peer destination");
    peers.ptr[d] = peer_C[d];
  }
  return dgemm_impl(st, M, N, K, alpha, A, lda, 1, B, ldb, 1, 0.0, C, ldc, ws, ws_bytes, tile_hint,
                    peers);
}

}  // namespace adccb
