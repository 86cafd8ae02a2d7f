// Shared helpers for the adccb CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace adccb {

// thread-local error message returned by adccb_last_error()
void set_error(const std::string& msg);
int fail(const std::string& msg);   // sets the error and returns a non-zero status

#define ADCCB_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return ::adccb::fail(std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    }                                                                                    \
  } while (0)

#define ADCCB_LAUNCH_OK()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      return ::adccb::fail(std::string("kernel launch: ") + cudaGetErrorString(_e));     \
    }                                                                                    \
  } while (0)

int sm_count();        // of the current device
int current_device();

// launch counter (bench.py reports it as gpu_launches)
void count_launch(int n = 1);

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

}  // namespace adccb
