// STUDY (round-2 preparation, not part of libadccb.so): the two bandwidth kernels around the INT8
// slice GEMMs of an Ozaki-type FP64 GEMM emulation.
//   slice:      X[r, :] (FP64)  ->  e[r] = floor(log2 max_k |X[r,k]|) + 1  and  s planes P_p[r, :] (int8) with
//               X[r,k] ~= 2^e[r] * sum_p 2^(-beta (p+1)) P_p[r,k]      (|P_p| < 2^beta, truncation)
//   recombine:  C[m,n] (+)= 2^(ea[m] + eb[n] - beta (d + 2)) * D[m,n]   for the INT32 sum D of all
//               slice products with p + q = d
// Same arithmetic as tools/studies/ozaki_int8_gpu_probe.py::slice_rows (powers of two and trunc
// only, hence bit-identical to the torch version).  Compiled here, not yet run on a GPU.
#include <cstdint>
#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ X, long long ldx, long long K,
                                                         long long Kpad, int beta, int s,
                                                         int8_t* __restrict__ planes, long long plane_stride,
                                                         int* __restrict__ expo) {
  __shared__ double red[8];
  __shared__ int e_sh;
  const long long r = blockIdx.x;
  const double* row = X + r * ldx;
  double m = 0.0;
  for (long long k = threadIdx.x; k < K; k += blockDim.x) m = fmax(m, fabs(row[k]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mm = red[0];
    for (int w = 1; w < 8; ++w) mm = fmax(mm, red[w]);
    int e = 0;
    if (mm > 0.0) frexp(mm, &e);             // mm = f * 2^e with 0.5 <= f < 1, so |X / 2^e| < 1 strictly
                                             // (an entry equal to 2^e would overflow the first int8 slice)
    e_sh = e;
    expo[r] = e;
  }
  __syncthreads();
  const double scale = ldexp(1.0, -e_sh);
  const double step = ldexp(1.0, beta);
  for (long long k = threadIdx.x; k < Kpad; k += blockDim.x) {
    double v = (k < K) ? row[k] * scale : 0.0;   // |v| < 1
    for (int p = 0; p < s; ++p) {
      v *= step;
      const double t = trunc(v);
      planes[p * plane_stride + r * Kpad + k] = (int8_t)t;
      v -= t;
    }
  }
}

__global__ void recombine_kernel(long long M, long long N, const int32_t* __restrict__ D,
                                 const int* __restrict__ ea, const int* __restrict__ eb, int shift,
                                 double beta_c, double* __restrict__ C, long long ldc) {
  const long long n_total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N, n = i % N;
    const double v = ldexp((double)D[i], ea[m] + eb[n] - shift);
    double* c = C + m * ldc + n;
    *c = (beta_c == 0.0) ? v : v + beta_c * *c;
  }
}

}  // namespace

extern "C" {

// planes: s * rows * Kpad int8 (plane-major); Kpad >= K, a multiple of 16; |beta| <= 7 with int8
int study_slice_rows(void* stream, long long rows, long long K, long long Kpad, const double* X, long long ldx,
                     int beta, int s, int8_t* planes, int* expo) {
  if (rows <= 0 || K <= 0 || Kpad < K || beta < 1 || beta > 7 || s < 1) return 1;
  slice_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(X, ldx, K, Kpad, beta, s, planes,
                                                                     rows * Kpad, expo);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// C = beta_c * C + 2^(ea[m] + eb[n] - shift) * D   (shift = beta * (d + 2) for the diagonal p + q = d)
int study_recombine(void* stream, long long M, long long N, const int32_t* D, const int* ea, const int* eb,
                    int shift, double beta_c, double* C, long long ldc) {
  if (M <= 0 || N <= 0) return 1;
  recombine_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(M, N, D, ea, eb, shift, beta_c, C, ldc);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // extern "C"
