// params.cu — see params.cuh.  M-step semantics: src/nip.c:2010-2071.
#include "params.cuh"

namespace nipgpu {
namespace {

__global__ void k_prior_flags(const double* prior, const int* prior_off, const int* vars, int n,
                              int* flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int v = vars[k];
  int any = 0;
  for (int i = prior_off[v]; i < prior_off[v + 1]; i++)
    if (prior[i] > 0) any = 1;
  flags[k] = any;
}

__global__ void k_apply_vector(double* table, int n, int stride, int card, const double* vec,
                               const int* flag) {
  if (flag && *flag == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    table[i] *= vec[(i / stride) % card];
}

__global__ void k_fill(double* a, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    a[i] = v;
}

// one thread per block of `card0` entries; left-to-right sum like the reference
__global__ void k_normalise_cpd(double* c, long long size, int card0) {
  const long long nb = size / card0;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb;
       b += (long long)gridDim.x * blockDim.x) {
    double* p = c + b * card0;
    double s = 0;
    for (int i = 0; i < card0; i++) s += p[i];
    if (s != 0)
      for (int i = 0; i < card0; i++) p[i] /= s;
  }
}

__global__ void k_init_potential(double* table, int n, const double* cpt, FamMap fm) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int f = 0;
    for (int k = 0; k < fm.n; k++) f += ((i / fm.cstride[k]) % fm.card[k]) * fm.fstride[k];
    table[i] *= cpt[f];
  }
}

__global__ void k_finish_estep(const double* acc, int groups, long long stride, long long n,
                               double pseudo, const double* ll, const int* status, int n_series,
                               double* counts) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long j = tid; j < n; j += (long long)gridDim.x * blockDim.x) {
    double s = pseudo;
    for (int g = 0; g < groups; g++) s += acc[g * stride + j];
    counts[j] = s;
  }
}

// tail of the accumulator: [n] = sum of the per-series log-likelihoods, [n+1] = any BAD_LUCK flag
// (one CTA, fixed order: the sum does not depend on how many series there are per thread block)
__global__ void k_estep_tail(const double* ll, const int* status, int n_series, double* tail) {
  __shared__ double red[40];
  double L = 0, bad = 0;
  for (int i = threadIdx.x; i < n_series; i += blockDim.x) { L += ll[i]; bad += status[i] ? 1.0 : 0.0; }
  L = block_sum(L, red);
  bad = block_sum(bad, red);
  if (threadIdx.x == 0) { tail[0] = L; tail[1] = bad != 0 ? 1.0 : 0.0; }
}

int blocks_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

}  // namespace

int prior_flags(const double* prior, const int* prior_off, const int* vars, int n, int* flags,
                cudaStream_t st) {
  if (n <= 0) return NIPGPU_OK;
  k_prior_flags<<<blocks_for(n, 128), 128, 0, st>>>(prior, prior_off, vars, n, flags);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int apply_vector(double* table, int n, int stride, int card, const double* vec, const int* flag,
                 cudaStream_t st) {
  k_apply_vector<<<blocks_for(n, 256), 256, 0, st>>>(table, n, stride, card, vec, flag);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int fill(double* a, long long n, double value, cudaStream_t st) {
  if (n <= 0) return NIPGPU_OK;
  k_fill<<<blocks_for(n, 256), 256, 0, st>>>(a, n, value);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int normalise_cpd(double* counts, long long size, int card0, cudaStream_t st) {
  k_normalise_cpd<<<blocks_for(size / card0, 128), 128, 0, st>>>(counts, size, card0);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int init_potential(double* table, int n, const double* cpt, const FamMap& fm, cudaStream_t st) {
  k_init_potential<<<blocks_for(n, 256), 256, 0, st>>>(table, n, cpt, fm);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

// flag[0] = 1 if some observation is >= the cardinality of its column's variable
__global__ void k_check_obs(const int* __restrict__ obs, long long n, int n_obs, const int* __restrict__ col_card,
                            int* flag) {
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
       x += (long long)gridDim.x * blockDim.x)
    if (obs[x] >= col_card[x % n_obs]) *flag = 1;
}

int check_obs(const int* obs, long long rows, int n_obs, const int* col_card, int* flag, cudaStream_t st) {
  if (rows <= 0 || n_obs <= 0) return NIPGPU_OK;
  k_check_obs<<<blocks_for(rows * n_obs, 256), 256, 0, st>>>(obs, rows * n_obs, n_obs, col_card, flag);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int finish_estep(const double* acc, int groups, long long stride, long long n, double pseudo,
                 const double* ll, const int* status, int n_series, double* counts, cudaStream_t st) {
  k_finish_estep<<<blocks_for(n, 256), 256, 0, st>>>(acc, groups, stride, n, pseudo, ll, status,
                                                    n_series, counts);
  NIPGPU_LAUNCHED();
  k_estep_tail<<<1, 512, 0, st>>>(ll, status, n_series, counts + n);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

}  // namespace nipgpu
