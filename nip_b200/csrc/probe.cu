// probe.cu — on-device measurement of the FP64 tensor (DMMA), FP64 vector (DFMA)
// and HBM copy rates, used as roofline denominators by bench.py.
#include "common.cuh"

namespace nipgpu {
namespace {

__global__ void __launch_bounds__(256) k_probe_dmma(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;  // keep the loop alive
}

__global__ void __launch_bounds__(256) k_probe_dfma(double* out, int iters) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_probe_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace
}  // namespace nipgpu

using namespace nipgpu;

extern "C" int nipgpu_probe_peaks(int device, double* dmma_tflops, double* dfma_tflops, double* copy_gbs) {
  if (int e = nipgpu_device_check(device)) return e;
  NIPGPU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NIPGPU_CUDA(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  NIPGPU_CUDA(cudaEventCreate(&e0));
  NIPGPU_CUDA(cudaEventCreate(&e1));
  double* d = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d, 64));
  float ms = 0;
  const int iters = 20000, blocks = sms * 8;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
    NIPGPU_CUDA(cudaEventRecord(e0));
    k_probe_dmma<<<blocks, 256>>>(d, iters);
    NIPGPU_LAUNCHED();
    NIPGPU_CUDA(cudaEventRecord(e1));
    NIPGPU_CUDA(cudaEventSynchronize(e1));
    NIPGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * blocks;
    if (rep) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  if (dmma_tflops) *dmma_tflops = best;
  best = 0;
  for (int rep = 0; rep < 4; rep++) {
    NIPGPU_CUDA(cudaEventRecord(e0));
    k_probe_dfma<<<blocks, 256>>>(d, iters);
    NIPGPU_LAUNCHED();
    NIPGPU_CUDA(cudaEventRecord(e1));
    NIPGPU_CUDA(cudaEventSynchronize(e1));
    NIPGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * iters * 256.0 * blocks;
    if (rep) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  if (dfma_tflops) *dfma_tflops = best;
  const size_t bytes = (size_t)2 << 30;
  double2 *a = nullptr, *b = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&a, bytes));
  NIPGPU_CUDA(cudaMalloc((void**)&b, bytes));
  NIPGPU_CUDA(cudaMemset(a, 1, bytes));
  best = 0;
  for (int rep = 0; rep < 6; rep++) {
    NIPGPU_CUDA(cudaEventRecord(e0));
    k_probe_copy<<<sms * 16, 512>>>(a, b, bytes / sizeof(double2));
    NIPGPU_LAUNCHED();
    NIPGPU_CUDA(cudaEventRecord(e1));
    NIPGPU_CUDA(cudaEventSynchronize(e1));
    NIPGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
  }
  if (copy_gbs) *copy_gbs = best;
  cudaFree(a); cudaFree(b); cudaFree(d);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return NIPGPU_OK;
}

// ---- development probe: DMMA issue/latency behaviour of ONE warp per SM sub-partition.
// out[0] = clocks per DMMA with `chains` independent accumulators issued round-robin.
namespace nipgpu {
namespace {
template <int CH>
__global__ void k_probe_dmma_chain(double* out, long long* clk, int iters) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
}  // namespace
}  // namespace nipgpu

extern "C" int nipgpu_probe_dmma_chain(int chains, int warps_per_block, int blocks, double* clocks_per_dmma) {
  using namespace nipgpu;
  double* d = nullptr;
  long long* c = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d, 64));
  NIPGPU_CUDA(cudaMalloc((void**)&c, 64));
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) {
    switch (chains) {
      case 1: k_probe_dmma_chain<1><<<blocks, 32 * warps_per_block>>>(d, c, iters); break;
      case 2: k_probe_dmma_chain<2><<<blocks, 32 * warps_per_block>>>(d, c, iters); break;
      case 4: k_probe_dmma_chain<4><<<blocks, 32 * warps_per_block>>>(d, c, iters); break;
      case 8: k_probe_dmma_chain<8><<<blocks, 32 * warps_per_block>>>(d, c, iters); break;
      case 16: k_probe_dmma_chain<16><<<blocks, 32 * warps_per_block>>>(d, c, iters); break;
      default: return NIPGPU_EINVAL;
    }
    NIPGPU_LAUNCHED();
    NIPGPU_CUDA(cudaDeviceSynchronize());
  }
  long long h = 0;
  NIPGPU_CUDA(cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c);
  if (clocks_per_dmma) *clocks_per_dmma = (double)h / ((double)iters * chains);
  return NIPGPU_OK;
}
