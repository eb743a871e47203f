/* TEST INFRASTRUCTURE — CPU oracle for the NIP join-tree hot path.
 *
 * A plain-C restatement of the reference's algorithm, working on the same flat
 * model description the device library consumes (include/nipgpu.h).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
 * may load it; the product (nip_b200/) never does.
 *
 * Pinned against: the reference's own known answers (test/potentialtest.c,
 * test/cliquetest.c), the Appendix-C vectors of SURVEY.md, and side-by-side runs
 * of the reference itself (oracle/_ref/libnip_ref.so) — see tests/test_oracle*.py.
 */
#ifndef NIP_ORACLE_H
#define NIP_ORACLE_H

#include "nipgpu.h"

typedef struct orc_model orc_model;

/* ---- L1: potential algebra on bare arrays (src/nippotential.c) ---------- */
int orc_general_marginalise(const double* src, int src_ndim, const int* src_card,
                            double* dst, int dst_ndim, const int* dst_card, const int* mapping);
int orc_total_marginalise(const double* src, int ndim, const int* card, double* dst, int variable);
int orc_update_potential(const double* num, const double* den, int sub_ndim, const int* sub_card,
                         double* target, int ndim, const int* card, const int* mapping);
int orc_update_evidence(const double* num, const double* den, double* target, int ndim,
                        const int* card, int var);
int orc_init_potential(const double* probs, int sub_ndim, const int* sub_card, double* target,
                       int ndim, const int* card, const int* mapping);
void orc_normalise_array(double* a, int n);
void orc_normalise_cpd(double* a, int size, int card0);

/* ---- L2/L5 on a model --------------------------------------------------- */
orc_model* orc_model_new(const nipgpu_model_desc* d);
void orc_model_free(orc_model* m);

void orc_reset_model(orc_model* m);
void orc_total_reset(orc_model* m);
void orc_use_priors(orc_model* m, int has_history);
int orc_enter_evidence(orc_model* m, int var, const double* evidence);
int orc_enter_index_observation(orc_model* m, int var, int index);
void orc_make_consistent(orc_model* m);
double orc_prob_mass(orc_model* m);
int orc_marginal(orc_model* m, int var, double* out);
void orc_get_clique(orc_model* m, int clique, int original, double* out);
void orc_get_parameters(orc_model* m, double* clique_tables, double* var_prior);
void orc_set_parameters(orc_model* m, const double* clique_tables, const double* var_prior);

/* forward_inference / forward_backward_inference for one series */
int orc_infer(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
              const unsigned char* use_evidence, int n_query, const int* query_vars,
              int forward_only, int want_ll, double* post, double* loglik);

long orc_counts_size(const orc_model* m);
/* e_step for one series; counts accumulate (caller initialises, normally 1.0) */
int orc_estep(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
              const unsigned char* use_evidence, double* counts, double* loglik);
/* m_step; counts are normalised in place (normalise_cpd) */
int orc_mstep(orc_model* m, double* counts);

/* niplikelihood inner loop for one series: out[t] = {m1, m2} */
int orc_likelihood(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
                   const unsigned char* evidence_off, const unsigned char* evidence_on,
                   double* out);

#endif
