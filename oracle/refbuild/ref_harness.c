/* TEST INFRASTRUCTURE — never linked into the product.
 *
 * Thin C harness around the UNMODIFIED reference sources, compiled where they
 * lie (/root/reference/src) into oracle/_ref/libnip_ref.so.  It textually
 * includes src/nip.c so that the file-static e_step()/m_step()
 * (src/nip.c:1708, :2010) can be driven one iteration at a time for per-step
 * EM parity; everything else is reached through the public nip.h API.
 *
 * Used by tests/, bench.py's cpu_baseline / --impl reference legs and
 * tests/golden/make_golden.py (fixture generation) only.
 */
#define _GNU_SOURCE
#include "nip.c" /* resolved through -I/root/reference/src */

#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "nip_model_export.h"

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void* refh_parse_model(const char* path) { return parse_model((char*)path); }
void refh_free_model(void* m) { free_model((nip_model)m); }

nipgpu_model_desc* refh_export(void* m) { return nipgpu_desc_from_model((nip_model)m); }
void refh_free_desc(nipgpu_model_desc* d) { nipgpu_desc_free(d); }

void refh_mark_all(void* m, int on) {
  nip_model model = (nip_model)m;
  int i;
  for (i = 0; i < model->num_of_vars; i++) {
    if (on) nip_mark_variable(model->variables[i]);
    else nip_unmark_variable(model->variables[i]);
  }
}

void refh_mark_var(void* m, int var, int on) {
  nip_model model = (nip_model)m;
  if (on) nip_mark_variable(model->variables[var]);
  else nip_unmark_variable(model->variables[var]);
}

/* Builds a time_series exactly as read_timeseries() lays it out
 * (src/nip.c:541-659) from already-indexed data: data[t*n_obs + k] is the
 * state index of obs_vars[k] at slice t, < 0 = missing. */
void* refh_new_timeseries(void* m, int n_obs, const int* obs_vars, int T, const int* data) {
  nip_model model = (nip_model)m;
  time_series ts = (time_series)calloc(1, sizeof(time_series_struct));
  int i, k, t, h = 0;
  ts->model = model;
  ts->length = T;
  ts->num_of_observed = n_obs;
  ts->num_of_hidden = model->num_of_vars - n_obs;
  ts->hidden = (nip_variable*)calloc((size_t)(ts->num_of_hidden > 0 ? ts->num_of_hidden : 1),
                                     sizeof(nip_variable));
  ts->observed = (nip_variable*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(nip_variable));
  for (k = 0; k < n_obs; k++) ts->observed[k] = model->variables[obs_vars[k]];
  for (i = 0; i < model->num_of_vars; i++) {
    int is_obs = 0;
    for (k = 0; k < n_obs; k++)
      if (obs_vars[k] == i) is_obs = 1;
    if (!is_obs) ts->hidden[h++] = model->variables[i];
  }
  ts->data = (int**)calloc((size_t)(T > 0 ? T : 1), sizeof(int*));
  for (t = 0; t < T; t++) {
    ts->data[t] = (int*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(int));
    for (k = 0; k < n_obs; k++) ts->data[t][k] = data[(size_t)t * n_obs + k];
  }
  return ts;
}

void refh_free_timeseries(void* ts) { free_timeseries((time_series)ts); }

/* forward_inference / forward_backward_inference for one series; posteriors
 * are flattened to post[t][concat of query vars]. */
int refh_infer(void* tsp, int nq, const int* qvars, int forward_only, int want_ll,
               double* post, double* ll) {
  time_series ts = (time_series)tsp;
  nip_model model = ts->model;
  nip_variable* vars = (nip_variable*)calloc((size_t)(nq > 0 ? nq : 1), sizeof(nip_variable));
  uncertain_series ucs;
  double L = 0;
  int i, t, s;
  size_t o = 0;
  for (i = 0; i < nq; i++) vars[i] = model->variables[qvars[i]];
  ucs = forward_only ? forward_inference(ts, vars, nq, want_ll ? &L : NULL)
                     : forward_backward_inference(ts, vars, nq, want_ll ? &L : NULL);
  free(vars);
  if (!ucs) return NIP_ERROR_GENERAL;
  if (post)
    for (t = 0; t < ucs->length; t++)
      for (i = 0; i < nq; i++)
        for (s = 0; s < NIP_CARDINALITY(ucs->variables[i]); s++) post[o++] = ucs->data[t][i][s];
  if (ll && want_ll) *ll = L;
  free_uncertainseries(ucs);
  return NIP_NO_ERROR;
}

/* ---- EM, one step at a time ------------------------------------------- */
static nip_potential* new_parameters(nip_model model) {
  nip_potential* par = (nip_potential*)calloc((size_t)model->num_of_vars, sizeof(nip_potential));
  int v, i, n;
  for (v = 0; v < model->num_of_vars; v++) {
    int* card;
    n = nip_number_of_parents(model->variables[v]) + 1;
    card = (int*)calloc((size_t)n, sizeof(int));
    card[0] = NIP_CARDINALITY(model->variables[v]);
    for (i = 1; i < n; i++) card[i] = NIP_CARDINALITY(model->variables[v]->parents[i - 1]);
    par[v] = nip_new_potential(card, n, NULL);
    free(card);
  }
  return par;
}

static void free_parameters(nip_model model, nip_potential* par) {
  int v;
  for (v = 0; v < model->num_of_vars; v++) nip_free_potential(par[v]);
  free(par);
}

/* total size of the per-variable family tables (child first, then parents[]) */
long refh_counts_size(void* m) {
  nip_model model = (nip_model)m;
  nip_potential* par = new_parameters(model);
  long n = 0;
  int v;
  for (v = 0; v < model->num_of_vars; v++) n += par[v]->size_of_data;
  free_parameters(model, par);
  return n;
}

/* m_step (src/nip.c:2010) from a flat count vector; returns the normalised
 * CPTs in the same buffer. */
int refh_mstep(void* m, double* counts) {
  nip_model model = (nip_model)m;
  nip_potential* par = new_parameters(model);
  long o = 0;
  int v, i, e;
  for (v = 0; v < model->num_of_vars; v++)
    for (i = 0; i < par[v]->size_of_data; i++) par[v]->data[i] = counts[o++];
  e = m_step(par, model);
  o = 0;
  for (v = 0; v < model->num_of_vars; v++)
    for (i = 0; i < par[v]->size_of_data; i++) counts[o++] = par[v]->data[i];
  free_parameters(model, par);
  return e;
}

/* e_step (src/nip.c:1708) over a set of series with the model's current
 * parameters; counts start at 1.0 like em_learn does (src/nip.c:2171-2172). */
int refh_estep(void** tsp, int n_ts, double* counts, double* loglik) {
  time_series* set = (time_series*)tsp;
  nip_model model = set[0]->model;
  nip_potential* par = new_parameters(model);
  double probe = 0, L = 0;
  long o = 0;
  int v, i, n, e = NIP_NO_ERROR;
  for (v = 0; v < model->num_of_vars; v++) nip_uniform_potential(par[v], 1.0);
  for (n = 0; n < n_ts; n++) {
    e = e_step(set[n], par, &probe);
    if (e != NIP_NO_ERROR) break;
    L += probe;
  }
  for (v = 0; v < model->num_of_vars; v++)
    for (i = 0; i < par[v]->size_of_data; i++) counts[o++] = par[v]->data[i];
  *loglik = L;
  free_parameters(model, par);
  return e;
}

/* whole em_learn (src/nip.c:2076); the learning curve is copied out. */
int refh_em_learn(void** tsp, int n_ts, double threshold, long seed, double* curve,
                  int curve_cap, int* n_curve) {
  nip_double_list lc = nip_new_double_list();
  nip_double_link l;
  int e, k = 0;
  random_seed(&seed);
  e = em_learn((time_series*)tsp, n_ts, threshold, lc);
  for (l = NIP_LIST_ITERATOR(lc); l && k < curve_cap; l = NIP_LIST_NEXT(l)) curve[k++] = l->data;
  *n_curve = NIP_LIST_LENGTH(lc);
  nip_empty_double_list(lc);
  free(lc);
  return e;
}

/* rand()-based draw sequence em_learn would use for its initial parameters
 * (nip_random_potential, src/nippotential.c:222-229), in variable order. */
void refh_random_parameters(void* m, long seed, double* counts) {
  nip_model model = (nip_model)m;
  nip_potential* par = new_parameters(model);
  long o = 0;
  int v, i;
  random_seed(&seed);
  for (v = 0; v < model->num_of_vars; v++) {
    nip_random_potential(par[v]);
    for (i = 0; i < par[v]->size_of_data; i++) counts[o++] = par[v]->data[i];
  }
  free_parameters(model, par);
}

/* ---- niplikelihood inner loop (util/niplikelihood.c:111-135) ---------- */
int refh_likelihood(void* tsp, double* out /* [T][2] */) {
  time_series ts = (time_series)tsp;
  nip_model model = ts->model;
  int t;
  reset_model(model);
  use_priors(model, !NIP_HAD_A_PREVIOUS_TIMESLICE);
  for (t = 0; t < TIME_SERIES_LENGTH(ts); t++) {
    insert_ts_step(ts, t, model, NIP_MARK_OFF);
    make_consistent(model);
    out[2 * t] = model_prob_mass(model);
    insert_ts_step(ts, t, model, NIP_MARK_ON);
    make_consistent(model);
    out[2 * t + 1] = model_prob_mass(model);
    reset_model(model);
    use_priors(model, NIP_HAD_A_PREVIOUS_TIMESLICE);
  }
  return NIP_NO_ERROR;
}

/* ---- parameter access -------------------------------------------------- */
int refh_clique_size(void* m, int c) { return ((nip_model)m)->cliques[c]->p->size_of_data; }

void refh_get_clique(void* m, int c, int original, double* out) {
  nip_clique q = ((nip_model)m)->cliques[c];
  nip_potential p = original ? q->original_p : q->p;
  memcpy(out, p->data, sizeof(double) * (size_t)p->size_of_data);
}

void refh_get_prior(void* m, int var, double* out) {
  nip_variable v = ((nip_model)m)->variables[var];
  int i;
  for (i = 0; i < NIP_CARDINALITY(v); i++) out[i] = v->prior ? v->prior[i] : 0.0;
}

int refh_enter_evidence(void* m, int var, double* lik) {
  nip_model model = (nip_model)m;
  return nip_enter_evidence(model->variables, model->num_of_vars, model->cliques,
                            model->num_of_cliques, model->variables[var], lik);
}

int refh_marginal(void* m, int var, double* out) {
  nip_model model = (nip_model)m;
  double* p = get_probability(model, model->variables[var]);
  int i;
  if (!p) return NIP_ERROR_GENERAL;
  for (i = 0; i < NIP_CARDINALITY(model->variables[var]); i++) out[i] = p[i];
  free(p);
  return NIP_NO_ERROR;
}

/* generate_data's hand-over between two slices (src/nip.c:2433-2441, 2464-2474): alpha out of
 * the consistent tree, forget the evidence, priors with history, alpha into in_clique->p.  The
 * next make_consistent() has to start from THAT clique state. */
int refh_next_slice(void* m) {
  nip_model model = (nip_model)m;
  int n = model->outgoing_interface_size, i, e;
  int* card = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  nip_potential alpha;
  for (i = 0; i < n; i++) card[i] = NIP_CARDINALITY(model->outgoing_interface[i]);
  alpha = nip_new_potential(card, n, NULL);
  free(card);
  e = start_timeslice_message_pass(model, FORWARD, alpha);
  reset_model(model);
  use_priors(model, NIP_HAD_A_PREVIOUS_TIMESLICE);
  if (e == NIP_NO_ERROR) e = finish_timeslice_message_pass(model, FORWARD, alpha, NULL);
  nip_free_potential(alpha);
  return e;
}

/* sepset potentials in the description's order; which: 0 = old, 1 = new */
int refh_sepset_size(void* m, int s) {
  int ns = 0, n = -1;
  nip_sepset* seps = nipgpu_model_sepsets((nip_model)m, &ns);
  if (seps && s < ns) n = seps[s]->new->size_of_data;
  free(seps);
  return n;
}

void refh_get_sepset(void* m, int s, int which, double* out) {
  int ns = 0;
  nip_sepset* seps = nipgpu_model_sepsets((nip_model)m, &ns);
  if (seps && s < ns) {
    nip_potential p = which ? seps[s]->new : seps[s]->old;
    memcpy(out, p->data, sizeof(double) * (size_t)p->size_of_data);
  }
  free(seps);
}

/* ---- CPU baseline timing ------------------------------------------------
 * Smoothing over a set of series, optionally sharded over `nproc` forked
 * workers (the reference is single-threaded with global state, so processes
 * are the only way to use more cores).  Returns wall seconds. */
double refh_time_infer(void** tsp, int n_ts, int nq, const int* qvars, int want_ll, int nproc) {
  time_series* set = (time_series*)tsp;
  nip_model model = set[0]->model;
  nip_variable* vars = (nip_variable*)calloc((size_t)(nq > 0 ? nq : 1), sizeof(nip_variable));
  double t0, t1, L;
  int i, w;
  for (i = 0; i < nq; i++) vars[i] = model->variables[qvars[i]];
  t0 = now_s();
  if (nproc <= 1) {
    for (i = 0; i < n_ts; i++) {
      uncertain_series u = forward_backward_inference(set[i], vars, nq, want_ll ? &L : NULL);
      free_uncertainseries(u);
    }
  } else {
    for (w = 0; w < nproc; w++) {
      pid_t pid = fork();
      if (pid == 0) {
        for (i = w; i < n_ts; i += nproc) {
          uncertain_series u = forward_backward_inference(set[i], vars, nq, want_ll ? &L : NULL);
          free_uncertainseries(u);
        }
        _exit(0);
      }
    }
    for (w = 0; w < nproc; w++) wait(NULL);
  }
  t1 = now_s();
  free(vars);
  return t1 - t0;
}

/* One EM iteration (M-step from `counts`, then E-step) timed on one core. */
double refh_time_em_iteration(void** tsp, int n_ts, double* counts, double* loglik) {
  double t0 = now_s();
  refh_mstep(((time_series*)tsp)[0]->model, counts);
  refh_estep(tsp, n_ts, counts, loglik);
  return now_s() - t0;
}

/* ---- helpers for tests of the nip.h drop-in (nip_b200/host/nip_gpu_backend.c) ---- */
void* refh_variable(void* m, int var) { return ((nip_model)m)->variables[var]; }

int refh_ucs_length(void* u) { return ((uncertain_series)u)->length; }

void refh_flatten_ucs(void* up, double* out) {
  uncertain_series u = (uncertain_series)up;
  int t, i, s;
  size_t o = 0;
  for (t = 0; t < u->length; t++)
    for (i = 0; i < u->num_of_vars; i++)
      for (s = 0; s < NIP_CARDINALITY(u->variables[i]); s++) out[o++] = u->data[t][i][s];
}

void refh_free_ucs(void* u) { free_uncertainseries((uncertain_series)u); }

void* refh_new_double_list(void) { return nip_new_double_list(); }

int refh_double_list_to_array(void* lp, double* out, int cap) {
  nip_double_list l = (nip_double_list)lp;
  nip_double_link k;
  int n = 0;
  for (k = NIP_LIST_ITERATOR(l); k && n < cap; k = NIP_LIST_NEXT(k)) out[n++] = k->data;
  return NIP_LIST_LENGTH(l);
}

void refh_free_double_list(void* lp) {
  nip_empty_double_list((nip_double_list)lp);
  free(lp);
}

long refh_seed(long seed) { return random_seed(&seed); }
