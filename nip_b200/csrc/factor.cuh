// factor.cuh — engine 3 (NIPGPU_ENGINE_FACTOR): the join tree evaluated FACTOR BY FACTOR.
//
// The reference materialises every clique table (nip_potential_struct, src/nippotential.h:46-60)
// and runs nip_message_pass (src/nipjointree.c:580-673) on them.  For models whose cliques hold
// millions of entries (config C3: 4 ring-coupled chains x 16 states -> three 16^6-entry cliques,
// 134 MB each) a slice-step then streams gigabytes.  A clique table, however, is by construction
// the product of the conditional probability tables of the families it hosts
// (nip_init_potential, src/nip.c:2044-2067), the evidence / prior vectors entered into it and the
// messages it absorbed.  This engine keeps that product SYMBOLIC: every message of the same join
// tree, and every marginal the slice loops of src/nip.c need, is one multi-operand contraction
//     out(O) = sum over R of  prod_k operand_k(vars_k)
// over the operands' own small tables (Shafer-Shenoy form: a message excludes what came from its
// receiver, so no division; the exact results are the ones the reference's Hugin form gives).
// C3: 16^6 multiply-adds per message instead of ~46 passes over 134 MB.
//
// What stays identical to the other engines: the alpha rows in HBM ([rows][S], one per slice),
// m1 = alpha_{t-1}.R1 and m2 = sum of the unnormalised alpha_t with the reference's
// log-likelihood / BAD_LUCK rules (src/nip.c:1458-1474, 1827-1831), alpha normalised per slice
// (zero sum: untouched), family marginals normalised by their own mass before they are summed
// into the expected counts (src/nip.c:1925-1967), the skip rule for previous-slice variables.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "model.h"

namespace nipgpu {

enum { FT_MODEL = 0, FT_SLOT = 1, FT_EVID = 2, FT_SAVED = 3 };   // SAVED: slot tensor kept per slice (the
                                                                 // upward messages, reused by the backward pass)

struct FacTensor {
  std::vector<int> vars;   // dimension 0 first (fastest)
  int kind = FT_MODEL;
  long long off = 0;       // inside d_fac (MODEL) or inside a slot's area (SLOT)
  long long size = 1;
  int col = -1;            // EVID: data column carrying the observation
};

struct FacOpRef {
  int tensor;
  bool inv;                // use 1/x (x == 0 -> 0): the Hugin division for messages to leaf cliques
};

// device form of one operand of a contraction (or of its result)
struct FacOpDev {
  long long off;           // offset of the tensor (MODEL: in d_fac, SLOT: in the slot area)
  int kind, inv, col;
  int olo, ohi, roff;      // positions of its index tables in the int pool
  int toff[4];             // offsets of the register tile's four elements (all 0: the operand does not
                           // hold the tile's variables); a tile runs along one variable whose cardinality
                           // is a multiple of 4, or over 2 x 2 states of two even-cardinality variables
  int toff2[4];            // the same for the second tile of a 2-D register tile
};

constexpr int kFacMaxOps = 8;
constexpr int kFacMaxVar = 4;   // operands that vary along the register tile

// One contraction.  A thread owns a register tile of TJ (x TK) results: TJ = 4 along one output
// variable (or 2 x 2 states of two even-cardinality variables) chosen so that the heavy operands
// do not hold it and are loaded once per term of the tile; TK = 4 along a second such tile when
// no in-loop operand holds variables of both (the term is then the outer product of two
// operand slices).  Thread index -> (offset of every operand, offset of the result) through the
// olo/ohi tables; consecutive threads run along the variable that is fastest in the heavy
// operands' memory.
struct FacStepDev {
  int n_thr, F, R, Rc, n_chunks, TJ;
  int cpc, othr;           // results of few threads: a CTA covers `cpc` consecutive chunks, `othr` threads each
  int n_out;               // entries of the result tensor
  int nSh, nVar, nO;       // in-loop operands shared by the tile / varying along it; epilogue operands
  int TK, nVar2;           // 2-D tile: TK results along a second variable, in-loop operands varying along it only
  FacOpDev opR[kFacMaxOps], opO[kFacMaxOps];   // opR: shared ones first, then the first tile variable's, then the second's
  FacOpDev out;
};

enum { FI_CONTRACT = 0, FI_SETTLE_FWD, FI_BETA_NORM, FI_COUNT, FI_QUERY };

struct FacInstr {
  int kind = FI_CONTRACT;
  FacStepDev step{};       // FI_CONTRACT
  long long src_off = 0;   // FI_COUNT / FI_QUERY / FI_BETA_NORM: SLOT offset of the vector
  int n = 0;               // its length
  int var = -1;            // FI_COUNT: variable whose family it is; FI_QUERY: position in the query
  long long dst_off = 0;   // FI_COUNT: offset in the counts; FI_QUERY: offset in the output row
  bool only_t0 = false;    // FI_COUNT of a previous-slice variable: first slices only
  bool needs_history = false;  // FI_CONTRACT / FI_BETA_NORM of the message to slice t-1: t > 0 only
  double flops = 0;
  std::string note;        // NIPGPU_FACTOR_TRACE: result and operand variables
};

struct FacProgram {
  std::vector<FacInstr> fwd, bwd;
  long long slot_doubles = 0;      // per-slot work area without the saved region
  long long saved_doubles = 0;     // upward messages of one slice (kept per slice when memory allows)
  int n_ups = 0;                   // leading instructions of fwd / bwd that compute them
  long long o_alpha_in = 0, o_beta = 0, o_alpha_new = 0, o_bprev = 0, o_partial = 0;
  long long partial_doubles = 0;
  int n_obs = 0;
  std::vector<int> marked_cols;    // columns whose evidence is entered
  int* d_pool = nullptr;           // index tables
  int* d_marked = nullptr;         // [n_obs] 1 = entered
  double flops_fwd = 0, flops_bwd = 0;
};

struct FacEngine {
  bool ok = false;
  std::string why;
  // model factors
  std::vector<FacTensor> tensors;          // MODEL tensors first (fixed), program tensors appended per plan
  int n_model_tensors = 0;
  std::vector<int> cpt_tensor;             // per variable: its CPT tensor (nparents > 0) or -1
  std::vector<int> prior_tensor;           // per variable: its prior tensor (parentless, not previous-slice) or -1
  std::vector<int> kappa_tensor;           // per clique: 0-dimensional constant or -1
  std::vector<std::vector<FacOpRef>> local; // per clique: model factors hosted there
  int a0_tensor = -1;                      // product of the previous-slice variables' priors over I_{t-1}
  long long fac_total = 0;
  double* d_fac = nullptr;
  std::vector<double> h_fac;               // host copy of the extracted factors (create / set_parameters)
  // tree rooted at out_clique (or clique 0 without an interface)
  int root = 0;
  std::vector<int> parent, psep, preorder;
  std::vector<std::vector<int>> children;
  // cached programs, keyed by (evidence columns, query, counts)
  std::map<std::vector<int>, FacProgram> programs;
  // run-time buffers
  double* d_slots = nullptr;
  size_t slots_cap = 0;
  double* d_alpha_wave = nullptr;   // [slots][longest series][S] forward rows of the sequences in flight
  size_t alpha_cap = 0;
  double* d_ll_run = nullptr;   // [max slots] running log-likelihood
  int* d_bad_run = nullptr;
  int max_slots = 0;
};

// structure + factor extraction from the host tables; sets fe.ok / fe.why
void fac_build(const HostModel& hm, FacEngine& fe, double rel_tol = 1e-12);
// re-extract after nipgpu_model_set_parameters (host tables given), false when they do not factor
bool fac_extract(const HostModel& hm, FacEngine& fe, const double* tables, double rel_tol = 1e-12);
// upload h_fac / derive the factors from the device parameters:
//   from_counts != nullptr: CPTs are the normalised family counts of the M-step (device)
int fac_refresh(const HostModel& hm, FacEngine& fe, const double* d_prior, const int* d_prior_flags,
                const double* d_counts_or_null, cudaStream_t st);
void fac_free(FacEngine& fe);

struct FacRunArgs {
  int n_series, n_obs, t_max;
  long long rows;
  const std::vector<int>* len;         // host lengths
  const std::vector<long long>* row_off;
  const std::vector<int>* obs_vars;
  const uint8_t* use_evidence;         // per variable, or nullptr = all columns
  const int* d_obs;
  const long long* d_row_off;
  int n_query;
  const int* query;                    // host
  int post_row;                        // doubles per output row
  double* d_post;                      // or nullptr
  int want_ll, forward_only;
  double* d_alpha;                     // unused (the engine keeps the forward rows of a wave itself)
  double* d_ll;
  int* d_status;
  const double* d_R1;
  const double* d_m10;
  double* d_acc;                       // E-step: [slots][acc_stride] accumulators or nullptr
  long long acc_stride;
  int acc_slots;
};
// number of sequences the engine wants in flight (accumulator groups of the E-step)
int fac_slots(const HostModel& hm, const FacEngine& fe, int n_series);
int fac_run(const HostModel& hm, FacEngine& fe, const FacRunArgs& a, cudaStream_t st);

}  // namespace nipgpu
