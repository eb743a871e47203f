// chain.cuh — engine 2 (NIPGPU_ENGINE_CHAIN): models whose slice is one
// interface clique {I_{t-1}, I_t} with leaf cliques hanging off I_t (HMM-style
// DBNs: configs C1/C2/C5 of SURVEY §8, and C4 up to |I| = 128).
//
// For such a model the slice-to-slice messages of
// start/finish_timeslice_message_pass (src/nip.c:1031-1098) are
//     alpha_t = normalise( (alpha_{t-1} . A) * lambda_t )
//     gamma_t = normalise( alpha_{t-1} * (A . (lambda_t * gamma_{t+1} / alpha_t)) )
// with A the interface clique table and lambda_t the product of the leaf tables
// restricted to the slice's evidence.  Batched over sequences the contraction
// is a dense FP64 GEMM and runs on the tensor cores (DMMA, mma.sync m8n8k4):
// one warp owns 8 sequences for their whole length, keeps alpha/gamma in
// registers in the MMA accumulator layout, and never synchronises with any
// other warp.
#pragma once

#include <vector>

#include "common.cuh"
#include "model.h"

namespace nipgpu {

struct ChainLeafHost {
  int clique = -1;             // real leaf clique, or -1 for an interface-variable pseudo leaf
  int var = -1;                // pseudo leaf: the I_t variable it carries evidence for
  std::vector<int> free_vars;  // non-sepset variables of the leaf, clique order
  std::vector<int> cfg_stride; // per free var: stride of its code inside the config index
  int n_cfg = 1;               // prod(card + 1); code == card means "no evidence"
  int miss_cfg = 0;            // config index with every code == card
  int proj = -1;               // projection leaf -> its sepset (base/off)
  std::vector<int> ip_to_s;    // [S] interface state -> sepset entry
  long long lam_off = 0;       // offset of Lambda_l[n_cfg][SP] inside d_lam
};

struct ChainModel {
  bool ok = false;
  bool dense = false;          // |I| > 64: per-slice tiled DMMA GEMM (dense.cu) instead of warp-resident recursion
  int S = 0, SP = 0, NT = 0;   // interface size, padded size (8*NT; multiple of 128 when dense)
  int c0 = -1;
  std::vector<ChainLeafHost> leaves;  // real leaves first, then pseudo leaves
  int n_real = 0;
  std::vector<int> var_leaf, var_slot;  // per variable: leaf carrying its evidence / digit slot, or -1
  std::vector<int> ent_of;     // [S*S] (i_prev * S + i_cur) -> entry of the interface clique
  // device
  int* d_ent_of = nullptr;
  int *d_ent_im = nullptr, *d_ent_ip = nullptr;  // entry of the interface clique -> (previous, current) state
  int* d_ip_to_s = nullptr;                      // [n_real][S] state -> leaf sepset entry
  double *d_Bf1 = nullptr, *d_Bb1 = nullptr, *d_Bb0 = nullptr;  // fragment-ordered SPxSP
  double *d_phi0 = nullptr, *d_lam0 = nullptr, *d_R1 = nullptr, *d_colsum = nullptr; // [SP]
  double* d_As = nullptr;      // [S][S] plain transition table (one-tile interfaces, chain_small.cuh)
  double m1_0 = 1.0;           // mass of the evidence-free first slice
  double* d_lam = nullptr;     // all Lambda tables
  long long lam_total = 0;
  int* d_leaf_meta = nullptr;  // flattened per-leaf metadata for the refresh kernel: [n_free, card[], stride[]]
  std::vector<int> leaf_meta_off;      // per real leaf: offset inside d_leaf_meta
  long long* d_miss_rows = nullptr;    // per real leaf: row of its no-evidence Lambda inside d_lam
  long long param_version = 0; // bumped by chain_refresh: invalidates cached evidence tables
};

struct ChainBatch {
  bool ready = false;
  std::vector<int> order;          // sorted position -> series (length descending)
  std::vector<int> len_sorted;
  int* d_order = nullptr;
  int* d_len_sorted = nullptr;
  int* d_cfg = nullptr;            // [rows] combined evidence index per data row
  // thread-per-sequence kernels (chain_small.cuh): time-major offsets and evidence index
  long long* d_toff = nullptr;     // [t_max + 1] (slice, sequence) pairs before slice t, sequences sorted by length
  int* d_cfgT = nullptr;           // [rows]
  std::vector<int> cfgT_key;       // evidence plan d_cfgT was built for
  double* d_alpha = nullptr;       // [rows][SP], same row order as the API
  // scale bookkeeping of the warp-pair kernels: binary exponent carried by every stored forward
  // row, and per series the sum of its last forward row with that row's exponent
  int* d_fexp = nullptr;           // [rows]
  double* d_zc = nullptr;          // [n_series]
  int* d_zf = nullptr;             // [n_series]
  double* d_rn = nullptr;          // [rows] E-step: 1 / (forward row . beta row), from the bookkeeping
  double* d_lam_static = nullptr;  // [SP] product of the inactive real leaves' no-evidence rows
  double* d_comb = nullptr;        // [n_comb][SP] combined evidence table of the cached plan
  size_t comb_cap = 0;
  int* d_cols = nullptr;           // column metadata of the cached plan
  long long* d_rows = nullptr;
  std::vector<int> plan_key;       // identifies the cached plan (evidence columns -> leaves)
  long long plan_version = -1;     // parameter version its evidence tables were built from
  int n_inactive = 0;
  // dense engine (dense.cu): per-sequence state, beta / R rows
  double* d_dense = nullptr;
  int* d_dense_i = nullptr;
  cudaStream_t dense_stream[3] = {};     // the other parts of the batch run here
  cudaEvent_t dense_fork = nullptr, dense_join[3] = {};
  // EM (chain_estep)
  double *d_rt = nullptr, *d_hvec = nullptr, *d_r0 = nullptr, *d_em_scratch = nullptr;  // beta rows, h_t, r_0 rows
  unsigned char* d_first = nullptr;  // [rows] 1 on the first row of every series
  size_t em_scratch_cap = 0;
};

// per-call evidence plan: which leaves see evidence through which data columns
struct ChainPlan {
  int n_active = 0;
  std::vector<int> active_leaf;       // leaf ids
  std::vector<int> col_leaf_slot;     // [n_obs] index into active_leaf or -1
  std::vector<int> col_stride;        // [n_obs] cfg stride of the column's digit
  std::vector<int> col_card;          // [n_obs]
  std::vector<int> col_mult;          // [n_obs] multiplier of the column's leaf in the combined index
  std::vector<int> mult;              // per active leaf
  int n_comb = 1, c_miss = 0;         // combined evidence configurations / the "no evidence" one
};

std::string chain_build(HostModel& hm, ChainModel& cm);
int chain_upload_structure(const HostModel& hm, ChainModel& cm, cudaStream_t st);
// recompute Bf1/Bb1/Bb0/phi0/lam0/Lambda from the base tables (after any parameter change)
int chain_refresh(const HostModel& hm, ChainModel& cm, const double* d_base0, const double* d_base1,
                  const std::vector<int>& tab_off, const int* d_ipool, const double* d_R1,
                  const double* d_m10, cudaStream_t st);
void chain_free(ChainModel& cm);

struct ChainInferArgs {
  int n_series, n_obs, t_max;
  long long rows;
  const int* d_obs;            // series-major [rows][n_obs]
  const long long* d_row_off;  // series-major first row per series
  int want_ll, forward_only;
  double* d_post;              // series-major rows, `post_stride` doubles each, or nullptr
  int post_stride, post_off;
  double* d_ll;                // [n_series] or nullptr
  int* d_status;               // [n_series] or nullptr
  int p0 = 0, p1 = -1;         // sorted positions [p0, p1) to process (p1 < 0: all) — chunked host copies
};

// returns false when the plan cannot be served by the chain engine
bool chain_plan(const HostModel& hm, const ChainModel& cm, int n_obs, const int* obs_vars,
                const uint8_t* use_evidence, ChainPlan& plan);
int chain_batch_prepare(const ChainModel& cm, ChainBatch& cb, int n_series, const int* len,
                        long long rows, int t_max, cudaStream_t st);
int chain_infer(const HostModel& hm, const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan,
                const ChainInferArgs& a, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1);
void chain_batch_free(ChainBatch& cb);
extern cudaEvent_t g_chain_mid_event;   // instrumentation: recorded between the forward and backward kernels
// dense.cu
int dense_refresh_mats(const ChainModel& cm, const double* d_base1_c0, cudaStream_t st);
// E-step outputs of the dense backward pass (see k_chain_stats in chain.cu for what they mean)
struct DenseEm {
  double* bt;     // [rows][SP] scaled beta_t
  double* hvec;   // [rows] h_t with beta_{t-1} = h_t A r_t
  double* r0;     // [n_series][SP] r_0 / (phi0 . r_0)
};
int dense_infer(const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan, const ChainInferArgs& a,
                cudaStream_t st, const DenseEm* em = nullptr);
// Statistics of the E-step for |I| > 64: per-row norms, posterior rows into evidence-indexed
// tables (partC: partsC x [n_comb][SP]) and the transition counts as a split-K TN DMMA GEMM
// (partG: partsG x [SP][SP]); `work` holds 2 x rows doubles.
int dense_stats(const ChainModel& cm, const ChainBatch& cb, const ChainPlan& plan, const DenseEm& em,
                const unsigned char* first, long long rows, double* work, int partsG, double* partG,
                int partsC, double* partC, cudaStream_t st);

struct ChainEmArgs {
  ChainInferArgs base;           // d_post / forward_only unused
  const double *d_base0, *d_base1;
  const std::vector<int>* tab_off;
  const int* d_ipool;
  double pseudo;                 // 1.0 = the reference's pseudo-count (src/nip.c:2171-2172), or 0
  double* d_counts;              // [counts + 2]
  int sm_count;
};
int chain_estep(const HostModel& hm, const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan,
                const ChainEmArgs& x, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1);

// ---- queries beyond the interface variables --------------------------------------------
// Every variable of a chain-structured model is a function of the posterior of the joint
// interface state gamma_t the engines produce:
//   kind 0  interface variable            marginal of gamma_t
//   kind 1  previous-slice interface var  marginal of gamma_{t-1}; on a series' first slice the
//           marginal of  sum_j gamma_0(j) base0(i, j) / phi0(j)             (smoothing only)
//   kind 2  free variable of a leaf       sum_s gamma_t(s) W[cfg_t][s][y],  W = leaf table
//           restricted to the slice's evidence / Lambda[cfg_t][s]
struct ChainQueryVar {
  int kind, var;
  int stride, card, off;         // joint-state stride and cardinality (kinds 0, 1); output offset
  int leaf, slot;                // kind 2
  int mult, n_cfg, fixed_cfg;    // kind 2: digit of the leaf inside the combined evidence index, or fixed
  long long w_off;               // kind 2: offset of W inside the scratch
};
// false when some queried variable is none of the three kinds (caller uses engine 1)
bool chain_query_plan(const HostModel& hm, const ChainModel& cm, const ChainPlan& plan, int nq,
                      const int32_t* query, int forward_only, std::vector<ChainQueryVar>& out);
int chain_post_vars(const HostModel& hm, const ChainModel& cm, const ChainBatch& cb,
                    const std::vector<ChainQueryVar>& qv, const double* d_base0, const double* d_base1,
                    const std::vector<int>& tab_off, const int* d_ipool, const double* joint,
                    const unsigned char* first, const long long* d_row_off, int n_series, long long rows,
                    int out_row, double* out, cudaStream_t st);

// Ancestral sampling of whole series from a chain-structured model (generate_data,
// src/nip.c:2325-2478, needs several make_consistent per variable per slice): one thread per
// series walks the slices — joint interface state from phi0 / the transition rows of base1, the
// previous-slice variables of slice 0 from base0, the free variables of every leaf from its table.
// out: device [n_series][length][n_vars] state indices, variables in model order.
int chain_sample(const HostModel& hm, const ChainModel& cm, const double* d_base0, const double* d_base1,
                 const std::vector<int>& tab_off, const int* d_ipool, int n_series, int length,
                 unsigned long long seed, int* d_out, cudaStream_t st);

}  // namespace nipgpu
