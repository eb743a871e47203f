/* nipgpu.h — C ABI of the B200 (sm_100a) backend for NIP's join-tree hot path.
 *
 * The reference (manuelschmidt/nip) has no plugin/FFI seam: its boundary is the
 * public C API of src/nip.h.  This header is what a `nip.h` implementation binds
 * to in order to run that API on the GPU:  plain pointers and sizes, no
 * reference types, no torch types.  Every entry point names the reference
 * function(s) it replaces.  The host-side glue that walks a `nip_model` and
 * calls these functions is nip_b200/host/nip_gpu_backend.c.
 *
 * Conventions
 *   - all functions return 0 on success or a NIPGPU_E* code; the message of the
 *     last failure on the calling thread is available from nipgpu_last_error();
 *   - all tables are IEEE double, flat, dimension 0 fastest
 *     (src/nippotential.c:58-68), indices are 32-bit like the reference's `int`;
 *   - variables are numbered 0..n_vars-1 in `model->variables[]` order
 *     (= ascending variable id = .net declaration order, src/nipvariable.c:72);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call
 *     fails with NIPGPU_ENODEVICE.
 */
#ifndef NIPGPU_H
#define NIPGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIPGPU_OK          0
#define NIPGPU_EINVAL      1   /* malformed description / argument            */
#define NIPGPU_ENODEVICE   2   /* no CUDA device / wrong architecture         */
#define NIPGPU_ECUDA       3   /* a CUDA runtime call failed                  */
#define NIPGPU_ENOMEM      4
#define NIPGPU_EUNSUPPORTED 5  /* model outside what the device path handles  */
#define NIPGPU_EBADLUCK    8   /* = NIP_ERROR_BAD_LUCK (src/nip.c:1827-1854)  */

/* interface_status bits, identical to src/nipvariable.h:30-34 */
#define NIPGPU_IF_INCOMING      1
#define NIPGPU_IF_OUTGOING      2
#define NIPGPU_IF_OLD_OUTGOING  4

/* engine selection for nipgpu_model_create (0 = let the compiler choose) */
#define NIPGPU_ENGINE_AUTO   0
#define NIPGPU_ENGINE_JTREE  1  /* generic join-tree schedule (any model)     */
#define NIPGPU_ENGINE_CHAIN  2  /* interface-clique + leaves, DMMA contraction*/
#define NIPGPU_ENGINE_FACTOR 3  /* join tree evaluated factor by factor: clique
                                   tables never materialised (huge cliques)   */

/* Flat description of one parsed time-slice model.  It is a by-value snapshot
 * of what parse_model() leaves in nip_model_struct (src/nip.h:71-112) and in
 * the nip_clique / nip_sepset / nip_variable structs reachable from it
 * (src/nipjointree.h:43-63, src/nipvariable.h:51-78).  The library copies
 * everything it needs; the caller keeps ownership of the arrays. */
typedef struct nipgpu_model_desc {
  /* ---- variables ---- */
  int32_t n_vars;
  const int32_t* var_card;        /* [n_vars] cardinality                        */
  const int32_t* var_flags;       /* [n_vars] NIPGPU_IF_* bits                   */
  const int32_t* var_parent_off;  /* [n_vars+1] offsets into var_parents         */
  const int32_t* var_parents;     /* parents in variable->parents[] order        */
  const int32_t* var_family;      /* [n_vars] family clique (nip_find_family)    */
  const int32_t* var_prior_off;   /* [n_vars+1] offsets into var_prior           */
  const double*  var_prior;       /* priors of parentless variables, 0 length
                                     for variables that have parents            */
  /* ---- cliques ---- */
  int32_t n_cliques;
  const int32_t* clique_var_off;  /* [n_cliques+1]                               */
  const int32_t* clique_vars;     /* variables of each clique, dim 0 first       */
  const int64_t* clique_tab_off;  /* [n_cliques+1] offsets into clique_tables    */
  const double*  clique_tables;   /* original_p of every clique, concatenated    */
  /* ---- sepsets ---- */
  int32_t n_sepsets;
  const int32_t* sepset_cliques;  /* [n_sepsets][2] first/second neighbour       */
  const int32_t* sepset_var_off;  /* [n_sepsets+1]                               */
  const int32_t* sepset_vars;     /* variables of each sepset, dim 0 first       */
  const int32_t* clique_adj_off;  /* [n_cliques+1] offsets into clique_adj       */
  const int32_t* clique_adj;      /* sepset ids in clique->sepsets list order    */
  /* ---- time-slice interface (src/nip.h:88-99) ---- */
  int32_t n_interface;            /* outgoing_interface_size                     */
  const int32_t* outgoing;        /* [n_interface] I_t      (alpha/gamma dims)   */
  const int32_t* prev_outgoing;   /* [n_interface] I_{t-1}, same index order     */
  int32_t in_clique;              /* -1 when n_interface == 0                    */
  int32_t out_clique;
} nipgpu_model_desc;

typedef struct nipgpu_model nipgpu_model;   /* compiled model, device resident  */
typedef struct nipgpu_batch nipgpu_batch;   /* a set of time series in HBM      */

const char* nipgpu_last_error(void);
/* 0 when a CUDA device of compute capability 10.x is usable */
int nipgpu_device_check(int device);

/* Compile a model: index maps / projections, message schedule, device upload
 * of original_p and priors.  Replaces the per-call nip_mapper()+malloc of
 * nip_message_pass (src/nipjointree.c:676-709) and the per-entry
 * nip_inverse_mapping loops (src/nippotential.c:251-264).
 * engine: NIPGPU_ENGINE_AUTO lets the library choose (chain-structured model ->
 * CHAIN; clique tables that fit shared memory -> JTREE; larger ones -> FACTOR
 * when every table verifiably is the product of its families' CPTs, which is
 * what parse_model / m_step build, src/nip.c:2044-2067; else JTREE streaming
 * them).  Asking for CHAIN or FACTOR on a model they cannot serve fails with
 * NIPGPU_EUNSUPPORTED. */
int nipgpu_model_create(const nipgpu_model_desc* desc, int device, int engine,
                        nipgpu_model** out);
void nipgpu_model_destroy(nipgpu_model* m);
/* which engine the compiler picked (NIPGPU_ENGINE_*) */
int nipgpu_model_engine(const nipgpu_model* m);
/* Host-only check, no device needed: is every clique table of the description the
 * product of the CPTs of the families it hosts (what parse_model and m_step build,
 * src/nip.c:2044-2067), i.e. can NIPGPU_ENGINE_FACTOR serve the model?  1: yes;
 * 0: no (the reason is in nipgpu_last_error()); < 0: malformed description. */
int nipgpu_model_factorable(const nipgpu_model_desc* desc);

/* Replace all clique tables / priors (same layout as in the description), e.g.
 * after the host changed original_p.  Replaces nothing in the reference: it is
 * the H2D half of keeping `clique->original_p` authoritative on the host. */
int nipgpu_model_set_parameters(nipgpu_model* m, const double* clique_tables,
                                const double* var_prior);
/* D2H: current device parameters → host arrays (after EM), so write_model()
 * (src/nip.c:298-482) sees the trained CPTs. */
int nipgpu_model_get_parameters(nipgpu_model* m, double* clique_tables,
                                double* var_prior);

/* Upload a set of time series (time_series_struct, src/nip.h:117-131).
 *   n_series    number of series
 *   lengths     [n_series] slices per series (ragged sets allowed)
 *   n_obs       number of observed columns (ts->num_of_observed)
 *   obs_vars    [n_obs] model variable of each column (ts->observed[])
 *   data        concatenated rows, series after series: sum(lengths) x n_obs
 *               state indices, < 0 = missing (src/nip.c:649, :994)
 * Replaces the per-slice insert_ts_step() evidence entry (src/nip.c:982-1001):
 * observations stay packed in HBM for the whole job. */
int nipgpu_batch_create(nipgpu_model* m, int32_t n_series, const int32_t* lengths,
                        int32_t n_obs, const int32_t* obs_vars, const int32_t* data,
                        nipgpu_batch** out);
void nipgpu_batch_destroy(nipgpu_batch* b);
/* Replace the observations of an existing batch (same shape as at creation):
 * the host-to-device half of one end-to-end step.  `data` may be pinned. */
int nipgpu_batch_update(nipgpu_batch* b, const int32_t* data);

/* Batched forward_inference / forward_backward_inference
 * (src/nip.c:1103-1315, 1320-1581) for every series of the batch.
 *   use_evidence [n_vars] non-zero = variable is marked (NIP_MARK_ON): only
 *                these columns enter evidence (src/nip.c:993); NULL = all
 *   query_vars   [n_query] variables of interest
 *   forward_only non-zero = filtering (forward_inference)
 *   post         host, sum(lengths) rows x (sum of card(query_vars)) doubles:
 *                row r of series s, variables in query order, or NULL
 *   loglik       host [n_series] total log-likelihood per series as the
 *                reference accumulates it (sum_t log m2 - log m1), or NULL
 * Host buffers; H2D of nothing (batch is resident), D2H of post/loglik.
 * Engine: chain-structured models (one clique over I_{t-1} and I_t, leaf cliques on I_t) run on
 * the tensor path for any query — interface, previous-slice interface and leaf variables are
 * computed from the posterior of the joint interface state; only filtering of a previous-slice
 * variable and evidence on one use the generic join-tree engine, as every other model does. */
int nipgpu_infer(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence,
                 int32_t n_query, const int32_t* query_vars, int forward_only,
                 double* post, double* loglik);

/* Same computation, results left in HBM (for callers that chain device work
 * and for kernel-only timing).  *post_dev / *loglik_dev receive device
 * pointers owned by the batch, valid until the next call on that batch. */
int nipgpu_infer_device(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence,
                        int32_t n_query, const int32_t* query_vars, int forward_only,
                        int want_loglik, double** post_dev, double** loglik_dev);

/* One E-step over the batch with the current device parameters
 * (e_step for every series, src/nip.c:1708-2007): expected family counts are
 * accumulated in HBM, starting from the reference's 1.0 pseudo-count
 * (src/nip.c:2171-2172) when `add_pseudocount` is non-zero.
 *   counts   host or NULL: per-variable family tables, child first then
 *            parents[] order (src/nip.c:2108-2128), concatenated in variable
 *            order; use nipgpu_model_counts_size() for the length
 *   loglik   sum over series
 *   status   0, or NIPGPU_EBADLUCK if any slice had m1<=0, m2<=0 or a running
 *            log-likelihood > 0 (src/nip.c:1827-1854) */
int nipgpu_em_estep(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence,
                    int add_pseudocount, double* counts, double* loglik, int* status);
int64_t nipgpu_model_counts_size(const nipgpu_model* m);
/* [n_vars+1] offsets of each variable's family table inside `counts` */
int nipgpu_model_counts_offsets(const nipgpu_model* m, int64_t* off);

/* Device pointer to the count accumulator (+2 trailing doubles: loglik, status)
 * so a multi-GPU caller can all-reduce it in place (one ncclAllReduce per EM
 * iteration) before the M-step. */
int nipgpu_em_counts_device(nipgpu_model* m, double** counts_dev, int64_t* n_doubles);

/* ---- several GPUs of one box ----------------------------------
 * A group is the same model compiled once per device (nipgpu_model_create with different
 * `device`).  The caller shards the series of a set over the members (one batch per member,
 * any rule: sequences are independent given the parameters, src/nip.c:2182-2207).  Inference
 * and the likelihood loop need no exchange: call nipgpu_infer on every member.  EM needs one:
 *   nipgpu_group_em_estep  every member runs e_step over its batch, then ONE
 *                          ncclAllReduce(sum, double) over [family counts | loglik | status]
 *                          leaves the set-wide totals in every member's HBM accumulator; the 1.0
 *                          pseudo-count (src/nip.c:2171-2172) enters once; counts / loglik /
 *                          status as in nipgpu_em_estep, reported from member 0
 *   nipgpu_group_em_mstep  m_step (src/nip.c:2010-2071) redundantly on every member, so the
 *                          parameters never leave HBM
 * A group of one member does no collective and does not need NCCL (bound at run time). */
typedef struct nipgpu_group nipgpu_group;
int nipgpu_group_create(nipgpu_model** models, int n, nipgpu_group** out);
void nipgpu_group_destroy(nipgpu_group* g);
int nipgpu_group_size(const nipgpu_group* g);
int nipgpu_group_em_estep(nipgpu_group* g, nipgpu_batch** batches, const uint8_t* use_evidence,
                          int add_pseudocount, double* counts, double* loglik, int* status);
int nipgpu_group_em_mstep(nipgpu_group* g, const double* counts);

/* M-step on the device from the accumulator (m_step, src/nip.c:2010-2071):
 * normalise_cpd, reset cliques to 1, multiply CPTs into family cliques,
 * priors <- normalised counts of parentless variables.  When `counts` is not
 * NULL it is uploaded first (used for the random initial parameters of
 * em_learn, src/nip.c:2135-2154). */
int nipgpu_em_mstep(nipgpu_model* m, const double* counts);

/* niplikelihood inner loop (util/niplikelihood.c:111-135) for every slice of
 * every series, slices evaluated independently (no inter-slice message):
 *   m1 = mass with the evidence of variables flagged in `evidence_off`
 *   m2 = mass with, in addition, the evidence of those in `evidence_on`
 * out: host, sum(lengths) x 2 doubles (m1, m2). */
int nipgpu_likelihood(nipgpu_model* m, nipgpu_batch* b, const uint8_t* evidence_off,
                      const uint8_t* evidence_on, double* out);

/* ---- single-slice, stateful API (B = 1) --------------------------------
 * Device counterparts of reset_model / use_priors / nip_enter_evidence /
 * make_consistent / model_prob_mass / get_probability (src/nip.c:61-119,
 * 1600-1617, 2254-2298; src/nipjointree.c:859-943) for callers that drive
 * slices by hand (util/nipjoint.c, test/hmmtest.c). */
int nipgpu_slice_reset(nipgpu_model* m);
int nipgpu_slice_use_priors(nipgpu_model* m, int has_history);
/* mark the prior of one parentless variable as entered (what use_priors does per variable) */
int nipgpu_slice_enter_prior(nipgpu_model* m, int32_t var);
int nipgpu_slice_enter_evidence(nipgpu_model* m, int32_t var, const double* likelihood);
int nipgpu_slice_make_consistent(nipgpu_model* m);
int nipgpu_slice_mass(nipgpu_model* m, double* mass);
int nipgpu_slice_marginal(nipgpu_model* m, int32_t var, double* out /*[card]*/);
/* D2H copy of a clique's current belief table (clique->p) */
int nipgpu_slice_get_clique(nipgpu_model* m, int32_t clique, double* out);
/* D2H copy of a sepset's current potential (sepset->new), sepsets numbered as in the description */
int nipgpu_slice_get_sepset(nipgpu_model* m, int32_t sepset, double* out);

/* make_consistent (src/nip.c:1600-1617) on the state the CALLER holds, whatever put it there
 * (reset_model/use_priors/nip_enter_evidence, but also finish_timeslice_message_pass,
 * src/nip.c:1069-1098, as generate_data does, :2433-2461): collect towards cliques[0], then
 * distribute, every message pass as nip_message_pass (src/nipjointree.c:676-709: swap old/new,
 * new = marginal of the sender, receiver *= new/old with x/0 -> 0).
 *   clique_tables   every clique->p, concatenated as in the description (clique_tab_off)
 *   sepset_tables   every sepset->new, concatenated in description order (may be NULL when the
 *                   model has no sepsets)
 *   clique_out      consistent clique->p, same layout (may alias clique_tables)
 *   sepset_new_out  sepset->new after the call (the distribute messages), or NULL
 *   sepset_old_out  sepset->old after the call (the collect messages), or NULL
 * One host-to-device copy, one kernel, one device-to-host copy per call (a captured CUDA graph
 * on persistent pinned/device buffers: no allocation after the first call). */
int nipgpu_slice_propagate(nipgpu_model* m, const double* clique_tables, const double* sepset_tables,
                           double* clique_out, double* sepset_new_out, double* sepset_old_out);

/* Ancestral sampling of n_series series of `length` slices from the model's current parameters
 * (generate_data, src/nip.c:2325-2478, one series at a time with several make_consistent per
 * variable per slice).  out: host [n_series][length][n_vars] state indices, variables in the
 * order of the description.  Chain-structured models only (NIPGPU_EUNSUPPORTED otherwise); the
 * random stream is the library's own (splitmix64 per series), not rand(). */
int nipgpu_sample(nipgpu_model* m, int32_t n_series, int32_t length, uint64_t seed, int32_t* out);

/* ---- instrumentation -------------------------------------------------- */
/* number of kernels this library launched since the counter was last reset */
int64_t nipgpu_launch_count(int reset);
/* milliseconds the GPU spent in the dominant kernel(s) of the last
 * nipgpu_infer / nipgpu_em_estep call, measured with CUDA events on the library's
 * stream; n receives the number of launches summed. */
int nipgpu_last_kernel_ms(nipgpu_model* m, double* ms, int32_t* n);
/* the forward kernel's share of the last nipgpu_infer on the tensor path (0 elsewhere) */
int nipgpu_last_forward_ms(nipgpu_model* m, double* ms);
/* Diagnostics of the generic engine's grid team (models whose cliques are streamed through
 * HBM; recording is on when the model was created with NIPGPU_JT_TRACE=1 in the environment):
 * copies up to cap_records (tag, device-timer ns) pairs, one per grid-wide barrier, into
 * out[2 * cap_records]; tag = operation code << 16 | projection.  Returns the number of records
 * (0 when recording is off), -1 on error. */
int nipgpu_jt_trace(nipgpu_model* m, uint64_t* out, int cap_records, int reset);
/* Measures, with CUDA events, what this device sustains on the two resources the
 * hot path is bound by (MEASURED_PEAKS.json has no FP64 figure): a dependent-free
 * stream of DMMA m8n8k4 instructions, the same with scalar DFMA, and a plain
 * device-to-device copy.  Used by bench.py as roofline denominators. */
int nipgpu_probe_peaks(int device, double* dmma_tflops, double* dfma_tflops, double* copy_gbs);
/* micro-probe behind the chain kernels' schedule: `chains` (1, 2, 4 or 8) independent
 * accumulator chains of DMMA m8n8k4 per warp, warps_per_block x blocks warps; reports the clocks
 * one warp spends per DMMA (issue interval when chains hide the latency, else the latency). */
int nipgpu_probe_dmma_chain(int chains, int warps_per_block, int blocks, double* clocks_per_dmma);
/* the chain kernels' sweep alone (128 DMMAs of one 8 x 64 . 64 x 64 contraction, B fragments from
 * shared memory), back to back on one warp per scheduler: clocks per sweep of B-fragment delivery
 * variant `variant` (nip_b200/csrc/sweep.cuh); 2048 is the tensor pipe's floor. */
int nipgpu_probe_sweep(int variant, int blocks, double* clocks_per_sweep);
/* scalar FP64 arithmetic and DMMA on one pipe?  clocks per DMMA of a warp that issues
 * `dfma_per_dmma` (0, 1, 2, 4) independent DFMAs after every DMMA */
int nipgpu_probe_dmma_dfma(int dfma_per_dmma, int warps_per_block, int blocks, double* clocks_per_dmma);
/* the CUDA stream (cudaStream_t) all work of this model is enqueued on */
void* nipgpu_model_stream(nipgpu_model* m);

#ifdef __cplusplus
}
#endif
#endif /* NIPGPU_H */
