/* nip_gpu_backend.c — the hot entry points of nip.h on the B200 backend.
 *
 * Drop-in replacements, with the reference's exact signatures and error
 * behaviour, for
 *     forward_inference            src/nip.c:1103-1315
 *     forward_backward_inference   src/nip.c:1320-1581
 *     em_learn                     src/nip.c:2076-2250   (e_step :1708, m_step :2010)
 *     make_consistent              src/nip.c:1600-1617
 * implemented as thin C glue over the C ABI of include/nipgpu.h.  A maintainer
 * links this file (and libnipgpu.so) into libnip INSTEAD of the four definitions
 * in src/nip.c; parser, join-tree construction, data readers, writers and every
 * util/ tool stay untouched (INTEGRATION.md shows the exact recipe).
 *
 * This file performs no potential arithmetic: it marshals `time_series` into the
 * packed int32 layout, keeps one compiled device model per `nip_model`, and
 * lays results out exactly as the callers free them (free_uncertainseries,
 * src/nip.c:896-908).  There is no CPU fallback: if the device library reports
 * an error the functions fail the way the reference fails (NULL / error code).
 *
 * Beyond nip.h it exports nip_gpu_smooth_set(), a batched variant for callers
 * that hold a whole `time_series` set (util/nipinference.c:125-132 loops over
 * one): same results, one device pass; and transparent batching for callers
 * that do NOT change: once a set is known (nip_gpu_register_set(), or the
 * read_timeseries wrapper below), the first per-series call smooths the whole
 * set in one pass and the following calls are served from that pass.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nip.h"
#include "nip_model_export.h"
#include "nipgpu.h"

#ifndef NIP_ERROR_GENERAL /* HEAD's niperrorhandler.h lost these; see oracle/refbuild */
#define NIP_ERROR_NULLPOINTER 1
#define NIP_ERROR_INVALID_ARGUMENT 3
#define NIP_ERROR_OUTOFMEMORY 4
#define NIP_ERROR_GENERAL 6
#define NIP_ERROR_BAD_LUCK 8
#endif
#define NIP_GPU_MIN_EM_ITERATIONS 3 /* MIN_EM_ITERATIONS, src/nip.c:29 */

/* ---- one compiled device model per host model ---------------------------- */
typedef struct {
  nip_model model;
  nipgpu_model_desc* desc;
  nipgpu_model* gm;
  double* tables; /* host parameters the device copy was built from */
  double* prior;
  long n_tables, n_prior;
  long version;   /* bumped whenever the parameters on the device change */
} backend_entry;

#define MAX_BACKENDS 16
static backend_entry backends[MAX_BACKENDS];
static int n_backends = 0;

static int nip_gpu_device(void) {
  const char* s = getenv("NIP_GPU_DEVICE");
  return s ? atoi(s) : 0;
}

static void gather_parameters(nip_model model, const nipgpu_model_desc* d, double* tables,
                              double* prior) {
  int i, j;
  for (i = 0; i < model->num_of_cliques; i++)
    memcpy(tables + d->clique_tab_off[i], model->cliques[i]->original_p->data,
           sizeof(double) * (size_t)model->cliques[i]->original_p->size_of_data);
  for (i = 0; i < model->num_of_vars; i++)
    if (model->variables[i]->num_of_parents == 0)
      for (j = 0; j < NIP_CARDINALITY(model->variables[i]); j++)
        prior[d->var_prior_off[i] + j] = model->variables[i]->prior ? model->variables[i]->prior[j] : 0.0;
}

/* Returns the device model of `model`, compiling it on first use and pushing
 * the host's original_p / priors again whenever they changed. */
static backend_entry* backend_for(nip_model model) {
  backend_entry* e = NULL;
  int i;
  for (i = 0; i < n_backends; i++)
    if (backends[i].model == model) e = &backends[i];
  if (!e) {
    if (n_backends == MAX_BACKENDS) return NULL;
    e = &backends[n_backends];
    memset(e, 0, sizeof(*e));
    e->desc = nipgpu_desc_from_model(model);
    if (!e->desc) return NULL;
    if (nipgpu_model_create(e->desc, nip_gpu_device(), NIPGPU_ENGINE_AUTO, &e->gm) != NIPGPU_OK) {
      fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
      nipgpu_desc_free(e->desc);
      return NULL;
    }
    e->model = model;
    e->n_tables = (long)e->desc->clique_tab_off[e->desc->n_cliques];
    e->n_prior = e->desc->var_prior_off[e->desc->n_vars];
    e->tables = (double*)malloc(sizeof(double) * (size_t)(e->n_tables > 0 ? e->n_tables : 1));
    e->prior = (double*)malloc(sizeof(double) * (size_t)(e->n_prior > 0 ? e->n_prior : 1));
    memcpy(e->tables, e->desc->clique_tables, sizeof(double) * (size_t)e->n_tables);
    memcpy(e->prior, e->desc->var_prior, sizeof(double) * (size_t)e->n_prior);
    n_backends++;
  }
  {
    double* t = (double*)malloc(sizeof(double) * (size_t)(e->n_tables > 0 ? e->n_tables : 1));
    double* p = (double*)malloc(sizeof(double) * (size_t)(e->n_prior > 0 ? e->n_prior : 1));
    gather_parameters(model, e->desc, t, p);
    if (memcmp(t, e->tables, sizeof(double) * (size_t)e->n_tables) ||
        memcmp(p, e->prior, sizeof(double) * (size_t)e->n_prior)) {
      if (nipgpu_model_set_parameters(e->gm, t, p) != NIPGPU_OK) {
        fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
        free(t); free(p);
        return NULL;
      }
      memcpy(e->tables, t, sizeof(double) * (size_t)e->n_tables);
      memcpy(e->prior, p, sizeof(double) * (size_t)e->n_prior);
      e->version++;
    }
    free(t); free(p);
  }
  return e;
}

/* Forget the device copy of a model (call before free_model()). */
static void forget_sets_of(nip_model model);

void nip_gpu_release(nip_model model) {
  int i;
  forget_sets_of(model);
  for (i = 0; i < n_backends; i++)
    if (backends[i].model == model) {
      nipgpu_model_destroy(backends[i].gm);
      nipgpu_desc_free(backends[i].desc);
      free(backends[i].tables); free(backends[i].prior);
      backends[i] = backends[--n_backends];
      return;
    }
}

/* ---- marshalling ---------------------------------------------------------- */
/* NIP_MARK_ON variables only enter evidence (insert_ts_step, src/nip.c:993) */
static unsigned char* marked_mask(nip_model model) {
  unsigned char* m = (unsigned char*)calloc((size_t)model->num_of_vars, 1);
  int i;
  if (m)
    for (i = 0; i < model->num_of_vars; i++) m[i] = (NIP_MARK(model->variables[i]) & NIP_MARK_ON) ? 1 : 0;
  return m;
}

static nipgpu_batch* upload_set(backend_entry* e, time_series* set, int n, long* rows_out) {
  nip_model model = e->model;
  int n_obs = set[0]->num_of_observed, s, t, k;
  long rows = 0, r = 0;
  int32_t *len, *vars, *data;
  nipgpu_batch* b = NULL;
  for (s = 0; s < n; s++) {
    if (set[s]->num_of_observed != n_obs) return NULL;
    rows += set[s]->length;
  }
  len = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  vars = (int32_t*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(int32_t));
  data = (int32_t*)calloc((size_t)(rows * n_obs > 0 ? rows * n_obs : 1), sizeof(int32_t));
  if (len && vars && data) {
    for (k = 0; k < n_obs; k++) vars[k] = nipgpu_var_index(model, set[0]->observed[k]);
    for (s = 0; s < n; s++) {
      len[s] = set[s]->length;
      for (t = 0; t < set[s]->length; t++)
        for (k = 0; k < n_obs; k++) data[r++] = set[s]->data[t][k];
    }
    if (nipgpu_batch_create(e->gm, n, len, n_obs, vars, data, &b) != NIPGPU_OK) {
      fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
      b = NULL;
    }
  }
  free(len); free(vars); free(data);
  if (rows_out) *rows_out = rows;
  return b;
}

/* the 3-level layout free_uncertainseries() expects (src/nip.c:896-908) */
static uncertain_series new_uncertain_series(nip_variable vars[], int nvars, int length,
                                             const double* flat, int row) {
  uncertain_series u = (uncertain_series)malloc(sizeof(uncertain_series_struct));
  int t, i, off;
  if (!u) return NULL;
  u->variables = (nip_variable*)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(nip_variable));
  u->data = (double***)calloc((size_t)(length > 0 ? length : 1), sizeof(double**));
  if (!u->variables || !u->data) { free(u->variables); free(u->data); free(u); return NULL; }
  memcpy(u->variables, vars, (size_t)nvars * sizeof(nip_variable));
  u->num_of_vars = nvars;
  u->length = length;
  for (t = 0; t < length; t++) {
    u->data[t] = (double**)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(double*));
    off = 0;
    for (i = 0; i < nvars; i++) {
      u->data[t][i] = (double*)calloc((size_t)NIP_CARDINALITY(vars[i]), sizeof(double));
      memcpy(u->data[t][i], flat + (size_t)t * row + off, sizeof(double) * (size_t)NIP_CARDINALITY(vars[i]));
      off += NIP_CARDINALITY(vars[i]);
    }
  }
  return u;
}

/* Smooths (or filters) a whole set in one device pass.  results[s] receives the
 * posterior series of set[s]; loglik, when not NULL, one value per series. */
int nip_gpu_smooth_set(time_series* set, int n, nip_variable vars[], int nvars, int forward_only,
                       uncertain_series* results, double* loglik) {
  backend_entry* e;
  nipgpu_batch* b;
  unsigned char* mask;
  int32_t* q;
  double* post;
  long rows = 0, r0 = 0;
  int i, s, row = 0, rc;
  if (!set || n <= 0 || !set[0] || !set[0]->model) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT, 1);
    return NIP_ERROR_INVALID_ARGUMENT;
  }
  e = backend_for(set[0]->model);
  if (!e) { nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1); return NIP_ERROR_GENERAL; }
  b = upload_set(e, set, n, &rows);
  mask = marked_mask(e->model);
  q = (int32_t*)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(int32_t));
  for (i = 0; i < nvars; i++) {
    q[i] = nipgpu_var_index(e->model, vars[i]);
    row += NIP_CARDINALITY(vars[i]);
  }
  post = (double*)calloc((size_t)(rows * row > 0 ? rows * row : 1), sizeof(double));
  if (!b || !mask || !q || !post) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY, 1);
    nipgpu_batch_destroy(b); free(mask); free(q); free(post);
    return NIP_ERROR_OUTOFMEMORY;
  }
  rc = nipgpu_infer(e->gm, b, mask, nvars, q, forward_only, nvars > 0 ? post : NULL, loglik);
  if (rc != NIPGPU_OK) {
    fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1);
  } else
    for (s = 0; s < n; s++) {
      results[s] = new_uncertain_series(vars, nvars, set[s]->length, post + (size_t)r0 * row, row);
      r0 += set[s]->length;
    }
  nipgpu_batch_destroy(b); free(mask); free(q); free(post);
  return rc == NIPGPU_OK ? NIP_NO_ERROR : NIP_ERROR_GENERAL;
}

/* ---- transparent batching -----------------------------
 * util/nipinference.c:125-129 and util/nipmap.c:139-145 call the smoother once per
 * series of a set they read with read_timeseries().  A registered set is smoothed
 * in ONE device pass on the first such call; the per-series results are parked
 * here and handed over (ownership included) as the caller asks for them.  A
 * parked pass is dropped when the query, the marked variables, the direction or
 * the model parameters differ from what it was computed for. */
typedef struct {
  time_series* set;        /* the caller's array (not owned) */
  int n;
  uncertain_series* res;   /* parked results, NULL once handed over */
  double* ll;
  int valid, forward_only, nvars;
  nip_variable* vars;
  unsigned char* mask;
  long version;
  nip_model model;
} set_entry;

#define MAX_SETS 16
static set_entry sets[MAX_SETS];
static int n_sets = 0;

static void drop_parked(set_entry* se) {
  int i;
  if (se->res)
    for (i = 0; i < se->n; i++)
      if (se->res[i]) { free_uncertainseries(se->res[i]); se->res[i] = NULL; }
  free(se->vars); free(se->mask);
  se->vars = NULL; se->mask = NULL;
  se->valid = 0;
}

/* Tell the backend that set[0..n-1] belong together (the array must stay alive until
 * nip_gpu_forget_set).  Hosts built with NIP_GPU_WRAP_SETS never call this: their
 * read_timeseries() does. */
void nip_gpu_register_set(time_series* set, int n) {
  set_entry* se;
  if (!set || n <= 1 || n_sets == MAX_SETS) return;
  se = &sets[n_sets];
  memset(se, 0, sizeof(*se));
  se->res = (uncertain_series*)calloc((size_t)n, sizeof(uncertain_series));
  se->ll = (double*)calloc((size_t)n, sizeof(double));
  if (!se->res || !se->ll) { free(se->res); free(se->ll); return; }
  se->set = set;
  se->n = n;
  n_sets++;
}

void nip_gpu_forget_set(time_series* set) {
  int i;
  for (i = 0; i < n_sets; i++)
    if (sets[i].set == set) {
      drop_parked(&sets[i]);
      free(sets[i].res); free(sets[i].ll);
      sets[i] = sets[--n_sets];
      return;
    }
}

/* a series is going away: the set it belongs to cannot be batched any more */
static void forget_series(time_series ts) {
  int i, k;
  for (i = 0; i < n_sets; i++)
    for (k = 0; k < sets[i].n; k++)
      if (sets[i].set[k] == ts) { nip_gpu_forget_set(sets[i].set); return; }
}

static uncertain_series batched_or_single(time_series ts, nip_variable vars[], int nvars, int forward_only,
                                          double* loglikelihood) {
  uncertain_series r = NULL;
  set_entry* se = NULL;
  int i, k, idx = -1;
  for (i = 0; i < n_sets && !se; i++)
    for (k = 0; k < sets[i].n; k++)
      if (sets[i].set[k] == ts) { se = &sets[i]; idx = k; break; }
  if (se && ts && ts->model) {
    backend_entry* e = backend_for(ts->model);
    unsigned char* mask = e ? marked_mask(e->model) : NULL;
    int same = e && mask && se->valid && se->model == ts->model && se->version == e->version &&
               se->forward_only == forward_only && se->nvars == nvars &&
               (nvars == 0 || memcmp(se->vars, vars, sizeof(nip_variable) * (size_t)nvars) == 0) &&
               memcmp(se->mask, mask, (size_t)e->model->num_of_vars) == 0;
    if (e && mask && !(same && se->res[idx])) {   /* nothing parked for this request: one pass over the set */
      drop_parked(se);
      se->vars = (nip_variable*)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(nip_variable));
      if (se->vars && nip_gpu_smooth_set(se->set, se->n, vars, nvars, forward_only, se->res, se->ll) == NIP_NO_ERROR) {
        memcpy(se->vars, vars, sizeof(nip_variable) * (size_t)nvars);
        se->mask = mask;
        mask = NULL;
        se->nvars = nvars; se->forward_only = forward_only; se->model = ts->model; se->version = e->version;
        se->valid = 1;
      } else {
        drop_parked(se);
      }
    }
    free(mask);
    if (se->valid && se->res[idx]) {
      r = se->res[idx];
      se->res[idx] = NULL;
      if (loglikelihood) *loglikelihood = se->ll[idx];
      return r;
    }
  }
  if (nip_gpu_smooth_set(&ts, 1, vars, nvars, forward_only, &r, loglikelihood) != NIP_NO_ERROR) return NULL;
  return r;
}

uncertain_series forward_backward_inference(time_series ts, nip_variable vars[], int nvars,
                                            double* loglikelihood) {
  return batched_or_single(ts, vars, nvars, 0, loglikelihood);
}

uncertain_series forward_inference(time_series ts, nip_variable vars[], int nvars,
                                   double* loglikelihood) {
  return batched_or_single(ts, vars, nvars, 1, loglikelihood);
}

static void forget_sets_of(nip_model model) {
  int i = 0;
  while (i < n_sets)
    if (sets[i].n > 0 && sets[i].set[0] && sets[i].set[0]->model == model) nip_gpu_forget_set(sets[i].set);
    else i++;
}

#ifdef NIP_GPU_WRAP_SETS
/* Built inside the reference tree, where src/nip.c is compiled with
 *   -Dread_timeseries=ref_read_timeseries -Dfree_timeseries=ref_free_timeseries
 * (INTEGRATION.md): every set the tools read is registered, every series they free is
 * forgotten, and the unchanged per-series loops run at batch throughput. */
int ref_read_timeseries(nip_model model, char* datafile, time_series** results);
void ref_free_timeseries(time_series ts);

int read_timeseries(nip_model model, char* datafile, time_series** results) {
  const int n = ref_read_timeseries(model, datafile, results);
  if (n > 1 && results && *results) nip_gpu_register_set(*results, n);
  return n;
}

void free_timeseries(time_series ts) {
  forget_series(ts);
  ref_free_timeseries(ts);
}
#endif

/* ---- data generation ---------------------------------------------------------
 * generate_data() (src/nip.c:2325-2478) samples one series with several make_consistent per
 * variable per slice.  nip_gpu_generate_set() draws a whole set on the device: n ordinary
 * `time_series` in which every model variable is observed (columns in model->variables order),
 * to be freed with free_timeseries().  The random stream is the device library's, not rand(). */
int nip_gpu_generate_set(nip_model model, int n, int length, unsigned long seed, time_series** results) {
  backend_entry* e;
  int32_t* raw;
  time_series* set;
  int s, t, k, nv, rc;
  if (!model || !results || n <= 0 || length <= 0) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT, 1);
    return 0;
  }
  *results = NULL;
  e = backend_for(model);
  if (!e) { nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1); return 0; }
  nv = model->num_of_vars;
  raw = (int32_t*)calloc((size_t)n * length * nv, sizeof(int32_t));
  set = (time_series*)calloc((size_t)n, sizeof(time_series));
  if (!raw || !set) { free(raw); free(set); nip_report_error(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY, 1); return 0; }
  rc = nipgpu_sample(e->gm, n, length, (uint64_t)seed, raw);
  if (rc != NIPGPU_OK) {
    fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1);
    free(raw); free(set);
    return 0;
  }
  for (s = 0; s < n; s++) {
    time_series ts = (time_series)calloc(1, sizeof(time_series_struct));
    ts->model = model;
    ts->length = length;
    ts->num_of_observed = nv;
    ts->num_of_hidden = 0;
    ts->observed = (nip_variable*)calloc((size_t)nv, sizeof(nip_variable));
    ts->hidden = (nip_variable*)calloc(1, sizeof(nip_variable));
    ts->data = (int**)calloc((size_t)length, sizeof(int*));
    for (k = 0; k < nv; k++) ts->observed[k] = model->variables[k];
    for (t = 0; t < length; t++) {
      ts->data[t] = (int*)calloc((size_t)nv, sizeof(int));
      for (k = 0; k < nv; k++) ts->data[t][k] = raw[((size_t)s * length + t) * nv + nipgpu_var_index(model, model->variables[k])];
    }
    set[s] = ts;
  }
  free(raw);
  *results = set;
  nip_gpu_register_set(set, n);
  return n;
}

/* ---- EM ---------------------------------------------------------------------
 * Control flow of em_learn (src/nip.c:2076-2250) with E- and M-steps on the
 * device: parameters, expected counts and log-likelihood stay in HBM across
 * iterations; only the scalar log-likelihood and the status come back per
 * iteration, the trained tables once at the end. */
int em_learn(time_series* ts, int n_ts, double threshold, nip_double_list learning_curve) {
  backend_entry* e;
  nipgpu_batch* b;
  nip_model model;
  unsigned char* mask;
  double *init, *tables, *prior;
  double old_ll, ll = -DBL_MAX;
  long n_counts, k = 0;
  int v, i, j, ts_steps = 0, status = 0, rc, iter = 0;
  if (!ts || n_ts <= 0 || !ts[0] || !ts[0]->model) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT, 1);
    return NIP_ERROR_INVALID_ARGUMENT;
  }
  model = ts[0]->model;
  if (learning_curve != NULL && NIP_LIST_LENGTH(learning_curve) > 0) nip_empty_double_list(learning_curve);
  e = backend_for(model);
  if (!e) { nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1); return NIP_ERROR_GENERAL; }
  b = upload_set(e, ts, n_ts, NULL);
  mask = marked_mask(model);
  n_counts = (long)nipgpu_model_counts_size(e->gm);
  init = (double*)malloc(sizeof(double) * (size_t)n_counts);
  if (!b || !mask || !init) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY, 1);
    nipgpu_batch_destroy(b); free(mask); free(init);
    return NIP_ERROR_OUTOFMEMORY;
  }
  /* random initial parameters: same rand() stream as nip_random_potential over
   * parameters[v] in variable order (src/nip.c:2135-2138, src/nippotential.c:222-229) */
  for (v = 0; v < model->num_of_vars; v++) {
    long size = NIP_CARDINALITY(model->variables[v]);
    for (j = 0; j < model->variables[v]->num_of_parents; j++)
      size *= NIP_CARDINALITY(model->variables[v]->parents[j]);
    for (i = 0; i < size; i++) init[k++] = rand() / (double)RAND_MAX;
  }
  for (i = 0; i < n_ts; i++) ts_steps += timeseries_length(ts[i]);

  rc = nipgpu_em_mstep(e->gm, init); /* first M-step enters the random parameters */
  free(init);
  for (;;) {
    if (rc != NIPGPU_OK) break;
    old_ll = ll;
    rc = nipgpu_em_estep(e->gm, b, mask, 1, NULL, &ll, &status);
    if (rc != NIPGPU_OK) break;
    if (status == NIPGPU_EBADLUCK) { rc = -NIP_ERROR_BAD_LUCK; break; } /* e_step's BAD_LUCK, :2185-2198 */
    if (learning_curve != NULL && nip_append_double(learning_curve, ll / ts_steps) != NIP_NO_ERROR) {
      rc = NIPGPU_ENOMEM;
      break;
    }
    if (old_ll > ll + ts_steps * threshold || ll > 0 || ll == -HUGE_DOUBLE) { /* :2224-2234 */
      rc = -NIP_ERROR_BAD_LUCK;
      break;
    }
    iter++;
    if (!((ll - old_ll) > ts_steps * threshold || iter < NIP_GPU_MIN_EM_ITERATIONS)) break; /* :2240-2241 */
    rc = nipgpu_em_mstep(e->gm, NULL); /* next M-step straight from the device accumulator */
  }
  /* the model keeps the parameters of the last M-step (src/nip.c:2149-2250): bring them home */
  tables = (double*)malloc(sizeof(double) * (size_t)(e->n_tables > 0 ? e->n_tables : 1));
  prior = (double*)malloc(sizeof(double) * (size_t)(e->n_prior > 0 ? e->n_prior : 1));
  if (tables && prior && nipgpu_model_get_parameters(e->gm, tables, prior) == NIPGPU_OK) {
    nipgpu_desc_store_parameters(model, e->desc, tables, prior);
    memcpy(e->tables, tables, sizeof(double) * (size_t)e->n_tables);
    memcpy(e->prior, prior, sizeof(double) * (size_t)e->n_prior);
  }
  free(tables); free(prior);
  nipgpu_batch_destroy(b); free(mask);
  if (rc == -NIP_ERROR_BAD_LUCK) return NIP_ERROR_BAD_LUCK;
  if (rc != NIPGPU_OK) {
    fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1);
    if (learning_curve != NULL) nip_empty_double_list(learning_curve);
    return NIP_ERROR_GENERAL;
  }
  return NIP_NO_ERROR;
}

/* ---- single slice -------------------------------------------------------------
 * make_consistent on the device, mirrored back so that the unchanged host code
 * (model_prob_mass, get_probability, nipjoint's direct reads of clique->p) sees a
 * consistent tree.  Evidence is taken from variable->likelihood, priors from
 * variable->prior_entered, exactly the state reset_model / use_priors /
 * nip_enter_evidence leave on the host. */
void make_consistent(nip_model model) {
  backend_entry* e = backend_for(model);
  int v, c, rc;
  if (!e) { nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1); return; }
  rc = nipgpu_slice_reset(e->gm);
  /* priors that use_priors() entered: flag them one by one through has_history */
  for (v = 0; rc == NIPGPU_OK && v < model->num_of_vars; v++)
    if (model->variables[v]->num_of_parents == 0 && model->variables[v]->prior_entered)
      rc = nipgpu_slice_enter_prior(e->gm, v);
  for (v = 0; rc == NIPGPU_OK && v < model->num_of_vars; v++)
    rc = nipgpu_slice_enter_evidence(e->gm, v, model->variables[v]->likelihood);
  if (rc == NIPGPU_OK) rc = nipgpu_slice_make_consistent(e->gm);
  for (c = 0; rc == NIPGPU_OK && c < model->num_of_cliques; c++)
    rc = nipgpu_slice_get_clique(e->gm, c, model->cliques[c]->p->data);
  if (rc == NIPGPU_OK) { /* sepsets: `new` is what nip_probability_mass reads (src/nipjointree.c:1146-1153) */
    int ns = 0, s;
    nip_sepset* seps = nipgpu_model_sepsets(model, &ns);
    for (s = 0; seps && rc == NIPGPU_OK && s < ns; s++) {
      rc = nipgpu_slice_get_sepset(e->gm, s, seps[s]->new->data);
      memcpy(seps[s]->old->data, seps[s]->new->data, sizeof(double) * (size_t)seps[s]->new->size_of_data);
    }
    free(seps);
  }
  if (rc != NIPGPU_OK) {
    fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error());
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1);
  }
}
