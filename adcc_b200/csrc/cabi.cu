// extern "C" surface of libadccb.so (see include/adccb.h for the contract).
#include "common.cuh"
#include "../../include/adccb.h"
#include <atomic>
#include <cstring>

namespace adccb {

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }
int fail(const std::string& msg) {
  g_error = msg;
  return 1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the CURRENT device (cached per device: a one-process-many-GPU host may mix parts)
int sm_count() {
  static std::atomic<int> cache[64];
  const int dev = current_device();
  int n = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int dgemm_strided(cudaStream_t st, long long M, long long N, long long K, double alpha,
                  const double* A, long long ars, long long acs, const double* B, long long brs,
                  long long bcs, double beta, double* C, long long ldc, double* ws,
                  long long ws_bytes, int tile_hint);
int dgemm_peers(cudaStream_t st, long long M, long long N, long long K, double alpha,
                const double* A, long long lda, const double* B, long long ldb, double* C,
                int n_peers, double* const* peer_C, long long ldc, double* ws, long long ws_bytes,
                int tile_hint);
int strided_gather(cudaStream_t st, int nd, const long long* shape, const long long* istride,
                   double alpha, const double* in, double gamma, const double* add, double beta,
                   double* out);
int elementwise(cudaStream_t st, int op, long long n, double alpha, const double* x,
                const double* y, double c, double beta, double* out);
int direct_sum(cudaStream_t st, long long na, long long nb, double sa, const double* a, double sb,
               const double* b, double* out);
int add_columns(cudaStream_t st, long long nrows, long long ncols, double* S, long long lds,
                long long col0, const double* G, long long ldg);
int lincomb(cudaStream_t st, int nvec, const double* coeff, const double* const* ptrs, long long n,
            double beta, double* out);
int lincomb_multi(cudaStream_t st, int nvec, int nout, const double* coeff, const double* const* ptrs,
                  long long n, double beta, double* const* outs);
int dot_many(cudaStream_t st, int nvec, const double* x, const double* const* ys, long long n,
             double scale, double beta, double* out_dev, double* ws, long long ws_bytes);
int spin_expand4(cudaStream_t st, const int* n, const double* T, double* out);
int spin_project(cudaStream_t st, int nd, const int* shape, const int* nalpha, double fac,
                 const double* in, double* out);
int spin_singlet(cudaStream_t st, const int* shape, const int* nalpha, double* T);
int precond_apply(cudaStream_t st, int nd, const int* shape, const int* nalpha, const double* x, const double* D,
                  double shift, int sym_mode, double spin_fac, int purify, double* out);

int packed_unpack(cudaStream_t st, int no, int nv, const double* U, double* out);
int packed_pack(cudaStream_t st, int no, int nv, const double* T, double* U);
int packed_expand(cudaStream_t st, int which, int no, int nv, const double* U, double* out);
int packed_expand_slab(cudaStream_t st, int which, int no, int nv, int first, int count, const double* U,
                       double* out);
int slab_layout(cudaStream_t st, int dir, long long nrows, long long npv, int world, const long long* cut,
                long long wmax, double* full, double* slabs);
int block_select(cudaStream_t st, int mode, long long nr, long long nc, const long long* rows,
                 const long long* cols, long long ld, double alpha, double* big, double* blk);
int vvvv_exchange_slab(cudaStream_t st, int nv, int a_begin, int na, const double* W, double* out);
int ovvv_exchange_slab(cudaStream_t st, int no, int nv, int a_begin, int na, const double* V2, double* out);
int pair_matvec(cudaStream_t st, int nv, int n_out, int n_red, long long so, long long sr, const double* V,
                const double* X, long long ldx, int x_by_out, double alpha, double beta, double* out,
                long long ldo, double* ws, long long ws_bytes);
int packed_fold_virt(cudaStream_t st, long long nrows, int nv, long long sa, const double* P1,
                     const long long* off1, const double* P2, const long long* off2,
                     const long long* rows, double alpha, double* S);
int packed_fold_occ(cudaStream_t st, int no, int nv, int j_begin, int nj, const double* C,
                    double alpha, double* S);
int packed_diagonal(cudaStream_t st, int no, int nv, const double* dov, const double* doo,
                    const double* dvv, double* D);
int pair_select(cudaStream_t st, long long L, int n, long long R, const double* in, double* out);
int synth_fill(cudaStream_t st, int layout, int blk, int no, int nv, unsigned long long seed,
               double scale, long long n, long long row_begin, int j_begin, int nj, double* out);

}  // namespace adccb

using namespace adccb;
static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");

extern "C" {

int adccb_version(void) { return 100; }
const char* adccb_last_error(void) { return g_error.c_str(); }
long long adccb_launch_count(void) { return g_launches.load(); }

int adccb_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  ADCCB_CUDA_OK(cudaGetDevice(&dev));
  ADCCB_CUDA_OK(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  ADCCB_CUDA_OK(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  ADCCB_CUDA_OK(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

int adccb_dgemm(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                int64_t a_rs, int64_t a_cs, const double* B, int64_t b_rs, int64_t b_cs,
                double beta, double* C, int64_t ldc, double* ws, int64_t ws_bytes, int tile_hint) {
  return dgemm_strided((cudaStream_t)stream, M, N, K, alpha, A, a_rs, a_cs, B, b_rs, b_cs, beta, C,
                       ldc, ws, ws_bytes, tile_hint);
}

int adccb_dgemm_peers(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                      int64_t lda, const double* B, int64_t ldb, double* C, int n_peers,
                      double* const* host_peer_C, int64_t ldc, double* ws, int64_t ws_bytes,
                      int tile_hint) {
  return dgemm_peers((cudaStream_t)stream, M, N, K, alpha, A, lda, B, ldb, C, n_peers, host_peer_C,
                     ldc, ws, ws_bytes, tile_hint);
}

int adccb_peer_alloc(int64_t bytes, void** dev_ptr, void* host_handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (bytes <= 0 || dev_ptr == nullptr || host_handle64 == nullptr) return fail("peer_alloc: bad arguments");
  void* p = nullptr;
  ADCCB_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(host_handle64, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

int adccb_peer_open(const void* host_handle64, void** dev_ptr) {
  if (dev_ptr == nullptr || host_handle64 == nullptr) return fail("peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle64, sizeof(h));
  void* p = nullptr;
  ADCCB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return 0;
}

int adccb_peer_close(void* dev_ptr) {
  ADCCB_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int adccb_peer_free(void* dev_ptr) {
  ADCCB_CUDA_OK(cudaFree(dev_ptr));
  return 0;
}

int adccb_strided_gather(void* stream, int ndim, const int64_t* shape, const int64_t* in_stride,
                         double alpha, const double* in, double gamma, const double* add,
                         double beta, double* out) {
  return strided_gather((cudaStream_t)stream, ndim, (const long long*)shape,
                        (const long long*)in_stride, alpha, in, gamma, add, beta, out);
}

int adccb_elementwise(void* stream, int op, int64_t n, double alpha, const double* x,
                      const double* y, double c, double beta, double* out) {
  return elementwise((cudaStream_t)stream, op, n, alpha, x, y, c, beta, out);
}

int adccb_direct_sum(void* stream, int64_t na, int64_t nb, double sa, const double* a, double sb,
                     const double* b, double* out) {
  return direct_sum((cudaStream_t)stream, na, nb, sa, a, sb, b, out);
}

int adccb_add_columns(void* stream, int64_t nrows, int64_t ncols, double* S, int64_t lds, int64_t col0,
                      const double* G, int64_t ldg) {
  return add_columns((cudaStream_t)stream, nrows, ncols, S, lds, col0, G, ldg);
}

int adccb_lincomb(void* stream, int nvec, const double* host_coeff, const double* const* host_ptrs,
                  int64_t n, double beta, double* out) {
  return lincomb((cudaStream_t)stream, nvec, host_coeff, host_ptrs, n, beta, out);
}

int adccb_lincomb_multi(void* stream, int nvec, int nout, const double* host_coeff,
                        const double* const* host_ptrs, int64_t n, double beta, double* const* host_outs) {
  return lincomb_multi((cudaStream_t)stream, nvec, nout, host_coeff, host_ptrs, n, beta, host_outs);
}

int adccb_dot_many(void* stream, int nvec, const double* x, const double* const* host_ys, int64_t n,
                   double scale, double beta, double* out_dev, double* ws, int64_t ws_bytes) {
  return dot_many((cudaStream_t)stream, nvec, x, host_ys, n, scale, beta, out_dev, ws, ws_bytes);
}

int adccb_spin_expand4(void* stream, const int* n_spatial, const double* T, double* out) {
  return spin_expand4((cudaStream_t)stream, n_spatial, T, out);
}

int adccb_spin_project(void* stream, int ndim, const int* shape, const int* n_alpha, double fac,
                       const double* in, double* out) {
  return spin_project((cudaStream_t)stream, ndim, shape, n_alpha, fac, in, out);
}

int adccb_spin_singlet(void* stream, const int* shape, const int* n_alpha, double* T) {
  return spin_singlet((cudaStream_t)stream, shape, n_alpha, T);
}

int adccb_precond_apply(void* stream, int ndim, const int* shape, const int* n_alpha, const double* x,
                        const double* D, double shift, int sym_mode, double spin_fac, int purify, double* out) {
  return precond_apply((cudaStream_t)stream, ndim, shape, n_alpha, x, D, shift, sym_mode, spin_fac, purify, out);
}

int adccb_packed_unpack(void* stream, int no, int nv, const double* U, double* out) {
  return packed_unpack((cudaStream_t)stream, no, nv, U, out);
}
int adccb_packed_pack(void* stream, int no, int nv, const double* T, double* U) {
  return packed_pack((cudaStream_t)stream, no, nv, T, U);
}
int adccb_packed_expand(void* stream, int which, int no, int nv, const double* U, double* out) {
  return packed_expand((cudaStream_t)stream, which, no, nv, U, out);
}
int adccb_packed_expand_slab(void* stream, int which, int no, int nv, int first, int count,
                              const double* U, double* out) {
  return packed_expand_slab((cudaStream_t)stream, which, no, nv, first, count, U, out);
}
int adccb_slab_layout(void* stream, int dir, int64_t nrows, int64_t npv, int world, const int64_t* host_cut,
                      int64_t wmax, double* full, double* slabs) {
  return slab_layout((cudaStream_t)stream, dir, nrows, npv, world, (const long long*)host_cut, wmax, full, slabs);
}
int adccb_copy2d(void* stream, void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                 int64_t width_bytes, int64_t height) {
  if (width_bytes <= 0 || height <= 0) return 0;
  ADCCB_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                                  (size_t)height, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}
int adccb_block_gather(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                       const double* in, int64_t ld, double* out) {
  return block_select((cudaStream_t)stream, 0, n_rows, n_cols, (const long long*)rows, (const long long*)cols, ld,
                      1.0, const_cast<double*>(in), out);
}
int adccb_block_scatter(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                        double alpha, const double* blk, double* out, int64_t ld) {
  return block_select((cudaStream_t)stream, 2, n_rows, n_cols, (const long long*)rows, (const long long*)cols, ld,
                      alpha, out, const_cast<double*>(blk));
}
int adccb_block_scatter_add(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                            double alpha, const double* blk, double* out, int64_t ld) {
  return block_select((cudaStream_t)stream, 1, n_rows, n_cols, (const long long*)rows, (const long long*)cols, ld,
                      alpha, out, const_cast<double*>(blk));
}
int adccb_vvvv_exchange_slab(void* stream, int nv, int a_begin, int na, const double* W, double* out) {
  return vvvv_exchange_slab((cudaStream_t)stream, nv, a_begin, na, W, out);
}
int adccb_ovvv_exchange_slab(void* stream, int no, int nv, int a_begin, int na, const double* V2, double* out) {
  return ovvv_exchange_slab((cudaStream_t)stream, no, nv, a_begin, na, V2, out);
}
int adccb_pair_matvec(void* stream, int nv, int n_out, int n_red, int64_t row_stride_out, int64_t row_stride_red,
                      const double* V, const double* X, int64_t ldx, int x_by_out, double alpha, double beta,
                      double* out, int64_t ldo, double* ws, int64_t ws_bytes) {
  return pair_matvec((cudaStream_t)stream, nv, n_out, n_red, row_stride_out, row_stride_red, V, X, ldx, x_by_out,
                     alpha, beta, out, ldo, ws, ws_bytes);
}
int adccb_packed_fold_virt(void* stream, int64_t nrows, int nv, int64_t sa, const double* P1,
                           const int64_t* off1, const double* P2, const int64_t* off2,
                           const int64_t* rows, double alpha, double* S) {
  return packed_fold_virt((cudaStream_t)stream, nrows, nv, sa, P1, (const long long*)off1, P2,
                          (const long long*)off2, (const long long*)rows, alpha, S);
}
int adccb_packed_fold_occ(void* stream, int no, int nv, int j_begin, int nj, const double* C,
                          double alpha, double* S) {
  return packed_fold_occ((cudaStream_t)stream, no, nv, j_begin, nj, C, alpha, S);
}
int adccb_packed_diagonal(void* stream, int no, int nv, const double* dov, const double* doo,
                          const double* dvv, double* D) {
  return packed_diagonal((cudaStream_t)stream, no, nv, dov, doo, dvv, D);
}
int adccb_pair_select(void* stream, int64_t L, int n, int64_t R, const double* in, double* out) {
  return pair_select((cudaStream_t)stream, L, n, R, in, out);
}
int adccb_synth_fill(void* stream, int layout, int blk, int no, int nv, uint64_t seed, double scale,
                     int64_t n, int64_t row_begin, int j_begin, int nj, double* out) {
  return synth_fill((cudaStream_t)stream, layout, blk, no, nv, seed, scale, n, row_begin, j_begin, nj, out);
}

}  // extern "C"
