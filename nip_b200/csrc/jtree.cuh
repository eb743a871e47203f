// jtree.cuh — generic join-tree engine (NIPGPU_ENGINE_JTREE): device program
// layout and launchers.  One CTA owns one sequence at a time and walks the
// compiled collect/distribute schedule for every slice with the sequence's
// clique tables staged in shared memory (or in a per-CTA HBM workspace when
// they do not fit); small models run one warp per sequence, huge cliques are
// streamed through HBM by the whole grid.
#pragma once

#include "common.cuh"

namespace nipgpu {

struct DProj {
  int tab;    // offset of the clique's table inside the work area
  int m, R, lanes;
  int base, off;  // positions in the int pool
  int clq;        // clique index
  int F, jlo, jhi;  // table-order map: j(e) = ipool[jhi + e / F] + ipool[jlo + e % F]
};

struct DMsg {
  int proj_src, proj_dst, slot, size;
};

// Everything a slice kernel needs about the model; passed by value.
struct DProgram {
  int tab_total, msg_total, msg_max, scratch;
  int n_cliques;
  int n_collect, n_distribute, n_path;
  int nif, S, proj_in, proj_out;
  int root_tab, root_size;
  int nv;
  const DProj* projs;
  const int* ipool;
  const DMsg* collect;
  const DMsg* distribute;
  const DMsg* path;
  const double* base0;  // original_p x every prior          (slices without history)
  const double* base1;  // original_p x priors of non-I_{t-1} variables (with history)
  const double* R1;     // [S] mass of base1 per interface state  (m1 = alpha . R1)
  const double* m1_0;   // [1] mass of base0                      (m1 at t = 0)
  const int* proj_var;  // [nv] family clique -> {v}
  const int* proj_fam;  // [nv] family clique -> (v, parents...)
  const long long* coff;  // [nv+1] family count offsets
  const int* var_flags;   // [nv]
};

struct DBatch {
  int n_series, n_obs;
  const int* len;          // [n_series]
  const long long* row_off;  // [n_series] first row of each series
  const int* obs;          // [rows][n_obs]
  const int* obs_proj;     // [n_obs] projection used to enter the column's evidence, -1 = ignored
};

struct DQuery {
  int n_query, row;        // row = doubles per output row
  const int* proj;         // [n_query]
  const int* off;          // [n_query] offset inside the row
};

// work-area size in doubles for one CTA
inline size_t jt_work_doubles(const DProgram& p, bool grid_team = false) {
  // the grid team also keeps a quotient vector (division-free distribute)
  const size_t quo = grid_team ? (size_t)(p.msg_max > p.S ? p.msg_max : p.S) : 0;
  return (size_t)p.tab_total + p.msg_total + p.msg_max + 3 * (size_t)p.S + p.scratch + quo + 40;
}

// Which set of threads owns one sequence (see jtree.cu).
enum { JT_MODE_CTA = 0, JT_MODE_WARP = 1, JT_MODE_GRID = 2 };

struct JtLaunch {
  int threads, grid;
  size_t smem_bytes;     // 0 when the work area lives in HBM
  double* gwork;         // per-CTA workspace (grid x work doubles), the grid team's single one, or nullptr
  int mode;              // JT_MODE_*
  int slots;             // teams running concurrently = accumulator groups of the E-step
  double* part;          // grid mode: 2 x grid doubles for grid-wide sums
  double* scratch;       // grid mode: grid x threads doubles for two-stage marginals
  unsigned long long* trace;  // grid mode, NIPGPU_JT_TRACE=1: per-barrier (tag, ns) records
  // grid mode with tables of a few MB: `groups` cooperative kernels run side by side, each with
  // grid CTAs, its own work area (group_stride doubles apart) and every groups-th sequence
  int groups;
  size_t group_stride;
  cudaStream_t aux_stream[8];
  cudaEvent_t aux_event[9];   // [0] fork, [g] join of group g
};

constexpr int JT_TRACE_WORDS = 1 + 2 * 8192;

// co-resident CTAs for the cooperative (grid-team) kernels
int jt_grid_ctas(int threads, int sm_count, size_t* smem_bytes);
// launch geometry for a batch of n_series sequences (never more teams than sequences)
JtLaunch jt_fit(const JtLaunch& l, int n_series);

int jt_forward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l,
               int want_ll, int emit_filtered, double* alpha, double* post, double* ll,
               int* status, cudaStream_t st);
int jt_backward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l,
                const double* alpha, double* post, double* acc /*per-CTA counts or null*/,
                long long acc_stride, cudaStream_t st);
int jt_likelihood(const DProgram& p, const DBatch& b, const int* proj_off, const int* proj_on,
                  const JtLaunch& l, double* out, cudaStream_t st);
// memoised likelihood loop: flags of the rows that open a series; per-record gather from the
// per-configuration table ([n_cfg][2 kinds][m1, m2])
int jt_first_rows(const long long* row_off, int n_series, long long rows, unsigned char* first, cudaStream_t st);
// host_stride0 / host_card0: the first column's stride and cardinality as the host knows them (a set
// with one data column takes a vectorised kernel)
int jt_like_gather(const int* obs, int n_obs, long long rows, const unsigned char* first, const int* col_stride,
                   const int* col_card, const double* table, double* out, int sm_count, cudaStream_t st,
                   int host_stride0 = 0, int host_card0 = 0);
int jt_calibrate(const DProgram& p, const JtLaunch& l, double* R1, double* m1_0, cudaStream_t st);
// single-slice propagation for the stateful API: tables <- base x lik, collect, distribute
int jt_slice(const DProgram& p, const JtLaunch& l, const double* start_tables, double* out_tables,
             double* out_msgs, cudaStream_t st);

// make_consistent on the caller's own state: in = [tables | sepsets], out = [tables | sepset new | sepset old]
int jt_propagate(const DProgram& p, const JtLaunch& l, const double* in, double* out, cudaStream_t st);

// reductions of the single-slice API on the consistent tables (device side)
int jt_mass(const double* tables, int n_tab, const double* msgs, int n_msg, double* out, cudaStream_t st);
int jt_marginal(const DProgram& p, const double* tables, int proj, double* out, cudaStream_t st);

}  // namespace nipgpu
