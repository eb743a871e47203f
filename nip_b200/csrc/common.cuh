// common.cuh — shared device helpers and host-side error plumbing.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "nipgpu.h"

namespace nipgpu {

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;  // kernels launched by this library (nipgpu_launch_count)

#define NIPGPU_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::nipgpu::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));        \
      return NIPGPU_ECUDA;                                                             \
    }                                                                                  \
  } while (0)

#define NIPGPU_LAUNCHED()                                                              \
  do {                                                                                 \
    ::nipgpu::g_launches++;                                                            \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::nipgpu::set_error(std::string("kernel launch: ") + cudaGetErrorString(e__));   \
      return NIPGPU_ECUDA;                                                             \
    }                                                                                  \
  } while (0)

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA in a fixed order (deterministic run to run); every
// thread gets the result.  `red` is 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = 0;
    for (int i = 0; i < nw; i++) s += red[i];
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}
#endif

}  // namespace nipgpu
