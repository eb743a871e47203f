// probe.cu — on-device measurement of the FP64 tensor (DMMA), FP64 vector (DFMA)
// and HBM copy rates, used as roofline denominators by bench.py.
#include "common.cuh"
#include "sweep.cuh"

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
// MODE 0: every DMMA reads the same A and B registers; 1: A differs per step of 8 DMMAs, B per
// DMMA (as in a k-step-major sweep); 2: A and B differ per DMMA
template <int CH, int MODE = 0>
__global__ void k_probe_dmma_chain(double* out, long long* clk, int iters) {
  double c[CH][2], av[16], bv[16];
#pragma unroll
  for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
#pragma unroll
  for (int i = 0; i < 16; i++) { av[i] = 1.0 + (threadIdx.x + i) * 1e-12; bv[i] = 1.0 - (threadIdx.x + 3 * i) * 1e-12; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < (MODE ? 16 / CH > 0 ? 16 / CH : 1 : 1); r++)
#pragma unroll
      for (int i = 0; i < CH; i++) {
        const int j = (r * CH + i) & 15;
        const double a = MODE == 0 ? av[0] : MODE == 1 ? av[r & 15] : av[j];
        const double b = MODE == 0 ? bv[0] : bv[j];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = (t1 - t0) / (MODE ? (16 / CH > 0 ? 16 / CH : 1) : 1);
}
}  // namespace
}  // namespace nipgpu

// ---- development probe: the chain kernels' sweep (sweep.cuh) alone, back to back, no side work.
// 128 DMMAs per sweep: 2048 clocks is the tensor pipe's floor for one warp per scheduler.
namespace nipgpu {
namespace {
template <int V>
__global__ void __launch_bounds__(256, 1) k_probe_sweep(double* out, long long* clk, int iters) {
  constexpr int NT = 8, SP = 64;
  extern __shared__ double sB[];
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) sB[i] = 1.0 / SP + 1e-9 * (i % 7);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double* frag = sB + 2 * lane;
  double own[NT][2], acc[NT][2];
#pragma unroll
  for (int n = 0; n < NT; n++) { own[n][0] = 1.0 + 1e-6 * lane; own[n][1] = 1.0 - 1e-6 * n; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    mma_sweep<NT, V>(acc, own, frag, [](auto) {});
#pragma unroll
    for (int n = 0; n < NT; n++) { own[n][0] = acc[n][0] * 0.015625; own[n][1] = acc[n][1] * 0.015625; }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int n = 0; n < NT; n++) s += own[n][0] + own[n][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
}  // namespace
}  // namespace nipgpu

// ---- does scalar FP64 arithmetic share the pipe the DMMAs run on?  4 accumulator chains, J
// independent DFMAs after every DMMA; one or two warps per scheduler.
namespace nipgpu {
namespace {
template <int J>
__global__ void k_probe_dmma_dfma(double* out, long long* clk, int iters) {
  double c[4][2], f[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = 1.0 + i * 1e-3;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int j = 0; j < J; j++)
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[(i * J + j) & 7]) : "d"(a), "d"(b));
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
}  // namespace
}  // namespace nipgpu

extern "C" int nipgpu_probe_dmma_dfma(int dfma_per_dmma, int warps_per_block, int blocks, double* clocks_per_dmma) {
  using namespace nipgpu;
  double* d = nullptr;
  long long* c = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d, 64));
  NIPGPU_CUDA(cudaMalloc((void**)&c, 64));
  const int iters = 2048;
  const dim3 g(blocks), b(32 * warps_per_block);
  for (int rep = 0; rep < 2; rep++)
    switch (dfma_per_dmma) {
      case 0: k_probe_dmma_dfma<0><<<g, b>>>(d, c, iters); break;
      case 1: k_probe_dmma_dfma<1><<<g, b>>>(d, c, iters); break;
      case 2: k_probe_dmma_dfma<2><<<g, b>>>(d, c, iters); break;
      case 4: k_probe_dmma_dfma<4><<<g, b>>>(d, c, iters); break;
      default: cudaFree(d); cudaFree(c); set_error("unsupported count"); return NIPGPU_EINVAL;
    }
  NIPGPU_LAUNCHED();
  long long h = 0;
  NIPGPU_CUDA(cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c);
  if (clocks_per_dmma) *clocks_per_dmma = (double)h / ((double)iters * 16);
  return NIPGPU_OK;
}

extern "C" int nipgpu_probe_sweep(int variant, int blocks, double* clocks_per_sweep) {
  using namespace nipgpu;
  double* d = nullptr;
  long long* c = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d, 64));
  NIPGPU_CUDA(cudaMalloc((void**)&c, 64));
  const int iters = 2000;
  const size_t smem = 64 * 64 * sizeof(double);
  const int threads = variant >= 10 ? 256 : 128;   // 10 + v: two warps per scheduler
  auto run = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) kernel<<<blocks, threads, smem>>>(d, c, iters);
  };
  switch (variant % 10) {
    case 0: run(k_probe_sweep<0>); break;
    case 1: run(k_probe_sweep<1>); break;
    case 2: run(k_probe_sweep<2>); break;
    case 3: run(k_probe_sweep<3>); break;
    default: cudaFree(d); cudaFree(c); set_error("unknown sweep variant"); return NIPGPU_EINVAL;
  }
  NIPGPU_LAUNCHED();
  long long h = 0;
  NIPGPU_CUDA(cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c);
  if (clocks_per_sweep) *clocks_per_sweep = (double)h / iters;
  return NIPGPU_OK;
}

extern "C" int nipgpu_probe_dmma_chain(int chains, int warps_per_block, int blocks, double* clocks_per_dmma) {
  using namespace nipgpu;
  double* d = nullptr;
  long long* c = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d, 64));
  NIPGPU_CUDA(cudaMalloc((void**)&c, 64));
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) {
    const dim3 g(blocks), b(32 * warps_per_block);
    switch (chains) {   // chains + 100 * operand mode
      case 1: k_probe_dmma_chain<1><<<g, b>>>(d, c, iters); break;
      case 2: k_probe_dmma_chain<2><<<g, b>>>(d, c, iters); break;
      case 4: k_probe_dmma_chain<4><<<g, b>>>(d, c, iters); break;
      case 8: k_probe_dmma_chain<8><<<g, b>>>(d, c, iters); break;
      case 16: k_probe_dmma_chain<16><<<g, b>>>(d, c, iters); break;
      case 102: k_probe_dmma_chain<2, 1><<<g, b>>>(d, c, iters); break;
      case 104: k_probe_dmma_chain<4, 1><<<g, b>>>(d, c, iters); break;
      case 108: k_probe_dmma_chain<8, 1><<<g, b>>>(d, c, iters); break;
      case 202: k_probe_dmma_chain<2, 2><<<g, b>>>(d, c, iters); break;
      case 204: k_probe_dmma_chain<4, 2><<<g, b>>>(d, c, iters); break;
      case 208: k_probe_dmma_chain<8, 2><<<g, b>>>(d, c, iters); break;
      default: cudaFree(d); cudaFree(c); set_error("unsupported chain count"); return NIPGPU_EINVAL;
    }
  }
  NIPGPU_LAUNCHED();
  long long h = 0;
  NIPGPU_CUDA(cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c);
  if (clocks_per_dmma) *clocks_per_dmma = (double)h / ((double)iters * (chains % 100));
  return NIPGPU_OK;
}
