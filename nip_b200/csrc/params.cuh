// params.cuh — small kernels that keep model parameters resident in HBM:
// base tables (original_p x priors), the M-step and count reductions.
#pragma once

#include "common.cuh"

namespace nipgpu {

constexpr int kMaxDims = 24;

// how a family table (child, parents...) sits inside its family clique
struct FamMap {
  int n;
  int cstride[kMaxDims];  // stride of the k-th family variable inside the clique table
  int card[kMaxDims];
  int fstride[kMaxDims];  // stride inside the family table (child fastest)
};

// flag[k] = any(prior_k > 0)          (nip_enter_prior refuses zero vectors,
//                                       src/nipjointree.c:917-927)
int prior_flags(const double* prior, const int* prior_off, const int* vars, int n, int* flags,
                cudaStream_t st);
// table[i] *= vec[(i / stride) % card]   when *flag != 0
int apply_vector(double* table, int n, int stride, int card, const double* vec, const int* flag,
                 cudaStream_t st);
int fill(double* a, long long n, double value, cudaStream_t st);
// nip_normalise_cpd (src/nippotential.c:373-383) on a family count table
int normalise_cpd(double* counts, long long size, int card0, cudaStream_t st);
// nip_init_potential (src/nippotential.c:525-564): table[i] *= cpt[family index of i]
int init_potential(double* table, int n, const double* cpt, const FamMap& fm, cudaStream_t st);
// counts[j] = pseudo + sum_g acc[g][j]  (fixed order); tail: [n] = sum ll, [n+1] = any status
int finish_estep(const double* acc, int groups, long long stride, long long n, double pseudo,
                 const double* ll, const int* status, int n_series, double* counts, cudaStream_t st);

// device-side range check of a batch's observations (flag[0] set when one is >= its cardinality)
int check_obs(const int* obs, long long rows, int n_obs, const int* col_card, int* flag, cudaStream_t st);

}  // namespace nipgpu
