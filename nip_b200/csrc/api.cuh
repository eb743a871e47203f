// api.cuh — the opaque objects behind include/nipgpu.h.
#pragma once

#include <vector>

#include "chain.cuh"
#include "common.cuh"
#include "factor.cuh"
#include "jtree.cuh"
#include "model.h"
#include "params.cuh"

struct nipgpu_model {
  nipgpu::HostModel hm;
  int device = 0;
  int engine = NIPGPU_ENGINE_JTREE;  // preferred engine
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
  cudaStream_t copy_stream = nullptr;      // chunked device-to-host copies of large posterior sets
  cudaEvent_t chunk_ev[4] = {};
  double last_kernel_ms = 0, last_forward_ms = 0;
  int last_kernel_n = 0;

  // ---- structure (immutable) ----
  int* d_ipool = nullptr;
  nipgpu::DProj* d_projs = nullptr;
  nipgpu::DMsg *d_collect = nullptr, *d_distribute = nullptr, *d_path = nullptr;
  int *d_proj_var = nullptr, *d_proj_fam = nullptr, *d_var_flags = nullptr;
  long long* d_coff = nullptr;
  int *d_prior_off = nullptr, *d_prior_vars = nullptr, *d_prior_flags = nullptr;
  std::vector<int> tab_off;  // per clique offset inside the table area (int)

  // ---- parameters and what is derived from them (HBM resident) ----
  double *d_orig = nullptr, *d_prior = nullptr;
  double *d_base0 = nullptr, *d_base1 = nullptr, *d_R1 = nullptr, *d_m10 = nullptr;
  double* d_counts = nullptr;  // family counts + {loglik, status}
  double* d_acc = nullptr;     // per-CTA accumulators
  size_t acc_groups = 0;
  double* d_gwork = nullptr;   // HBM workspace when tables do not fit shared memory
  size_t gwork_doubles = 0;
  unsigned long long* d_trace = nullptr;  // NIPGPU_JT_TRACE=1 (grid team diagnostics)
  cudaStream_t aux_stream[8] = {};        // grid team: concurrent groups
  cudaEvent_t aux_event[9] = {};

  nipgpu::DProgram prog{};
  nipgpu::JtLaunch launch{};

  nipgpu::ChainModel chain;  // engine 2 (valid when hm.chain_ok)
  nipgpu::FacEngine fac;     // engine 3 (valid when fac.ok)
  bool fac_auto = false;     // engine 3 was the library's own choice: a request it cannot plan goes to engine 1

  // ---- single-slice state (stateful API) ----
  std::vector<std::vector<double>> lik;  // host mirror of variable->likelihood
  std::vector<char> prior_entered;
  std::vector<double> lik_stage;         // [nv][card_max] staging of all evidence vectors
  double *d_slice_start = nullptr, *d_slice_tab = nullptr, *d_slice_msg = nullptr;
  bool slice_consistent = false;
  // nipgpu_slice_propagate: one pinned block [in | out], its device twin, and the captured
  // H2D -> kernel -> D2H graph (no allocation, one launch, one sync per call)
  double *h_prop = nullptr, *d_prop = nullptr;
  cudaGraphExec_t prop_graph = nullptr;
  double *d_vec = nullptr;   // card_max doubles: evidence vector of the stateful API
};

struct nipgpu_batch {
  nipgpu_model* m = nullptr;
  int n_series = 0, n_obs = 0, t_max = 0;
  long long rows = 0;
  std::vector<int> len, obs_vars;
  std::vector<long long> row_off;
  int* d_len = nullptr;
  long long* d_row_off = nullptr;
  int* d_obs = nullptr;
  int* d_obs_proj = nullptr;  // 3 x n_obs scratch (evidence / off / on sets)
  int *d_qproj = nullptr, *d_qoff = nullptr;
  size_t q_cap = 0;
  double *d_alpha = nullptr, *d_post = nullptr, *d_ll = nullptr, *d_like = nullptr;
  size_t post_cap = 0;
  int* d_status = nullptr;
  unsigned char* d_first = nullptr;  // [rows] 1 on the first row of every series (memoised likelihood)
  int* d_check = nullptr;            // [0] range-check flag, [1..n_obs] cardinality per data column
  double* d_joint = nullptr;         // [rows][SP] posterior of the joint interface state (composite interfaces)
  nipgpu::ChainBatch chain;
};

namespace nipgpu {
// the two halves of nipgpu_em_estep (api.cu); the multi-device group (group.cu) puts one
// all-reduce between them
int estep_enqueue(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int add_pseudocount);
int estep_finish(nipgpu_model* m, double* counts, double* loglik, int* status);
}  // namespace nipgpu
