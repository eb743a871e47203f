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
 * packed int32 layout, keeps one compiled device model per `nip_model` (and per
 * device of NIP_GPU_DEVICES), and lays results out exactly as the callers free
 * them (free_uncertainseries, src/nip.c:896-908).  There is no CPU fallback: if
 * the device library reports an error the functions fail the way the reference
 * fails (NULL / error code).
 *
 * make_consistent() works on the state the host model actually holds — every
 * clique->p and sepset potential goes to the device, collect/distribute run
 * there, the consistent tree comes back — so every caller of the reference is
 * served, including generate_data (src/nip.c:2433-2461), which multiplies the
 * inter-slice message into in_clique->p before calling it.
 *
 * Beyond nip.h it exports nip_gpu_smooth_set(), a batched variant for callers
 * that hold a whole `time_series` set (util/nipinference.c:125-132 loops over
 * one): same results, one device pass; and transparent batching for callers
 * that do NOT change: once a set is known (nip_gpu_register_set(), or the
 * read_timeseries wrapper below), the first per-series call smooths the whole
 * set in one pass and the following calls are served from that pass.
 *
 * Like the reference (parser statics, variable-id counter, rand()), this file
 * keeps process-global state and is NOT thread-safe: one thread per process
 * drives the nip.h API.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nip.h"
#include "nip_model_export.h"
#include "nipgpu.h"

#ifndef NIP_ERROR_GENERAL /* HEAD's niperrorhandler.h lost these; see nip_errcodes_shim.h */
#define NIP_ERROR_NULLPOINTER 1
#define NIP_ERROR_INVALID_ARGUMENT 3
#define NIP_ERROR_OUTOFMEMORY 4
#define NIP_ERROR_GENERAL 6
#define NIP_ERROR_BAD_LUCK 8
#endif
#define NIP_GPU_MIN_EM_ITERATIONS 3 /* MIN_EM_ITERATIONS, src/nip.c:29 */
#define NIP_GPU_MAX_DEVICES 8

static int report(const char* file, int line, int code) {
  nip_report_error(file, line, code, 1);
  return code;
}
static void report_device_error(void) { fprintf(stderr, "nip gpu backend: %s\n", nipgpu_last_error()); }

/* ---- memo of make_consistent (tiny models) ---------------------------------
 * util/niplikelihood.c:111-135 calls make_consistent twice per data record and the tree it
 * hands over takes only a handful of distinct values (one per evidence configuration, with
 * or without history).  A model whose whole state is a few hundred doubles keeps the device's
 * answers keyed by the exact bytes of the request: a repeated request costs a hash and a
 * memcmp instead of a launch and a synchronisation.  Every distinct state is still computed
 * on the device, once.  NIP_GPU_SLICE_MEMO=0 switches it off. */
typedef struct {
  unsigned long long hash;
  double* in;  /* n_in doubles followed by n_out doubles; NULL = empty slot */
} memo_slot;
#define MEMO_SLOTS 8192           /* power of two */
#define MEMO_MAX_FILL 4096
#define MEMO_MAX_DOUBLES 1024     /* request + answer */

/* ---- one compiled device model per host model ---------------------------- */
typedef struct {
  nip_model model;
  nipgpu_model_desc* desc;
  int n_dev;
  int dev[NIP_GPU_MAX_DEVICES];
  nipgpu_model* gm[NIP_GPU_MAX_DEVICES]; /* gm[0] serves inference and the single-slice API */
  nipgpu_group* group;                   /* n_dev > 1: EM across the devices */
  double* tables; /* host parameters the device copies were built from */
  double* prior;
  long n_tables, n_prior;
  long version;   /* bumped whenever the parameters on the device change */
  /* make_consistent: the model's sepsets in description order, staging [in | out] */
  nip_sepset* seps;
  int n_seps;
  long n_msg;
  double* stage;
  memo_slot* memo;
  int memo_fill, memo_on;
  unsigned long memo_hits, slice_calls;
} backend_entry;

static backend_entry* backends = NULL;
static int n_backends = 0, cap_backends = 0;

/* NIP_GPU_DEVICES=0,1,2,3 (EM shards the series over them) or NIP_GPU_DEVICE=k; default 0 */
static int configured_devices(int* dev) {
  const char* s = getenv("NIP_GPU_DEVICES");
  int n = 0;
  if (s && *s) {
    while (*s && n < NIP_GPU_MAX_DEVICES) {
      char* end = NULL;
      long v = strtol(s, &end, 10);
      if (end == s) break;
      dev[n++] = (int)v;
      s = end;
      while (*s == ',' || *s == ' ') s++;
    }
  }
  if (n == 0) {
    s = getenv("NIP_GPU_DEVICE");
    dev[n++] = s ? atoi(s) : 0;
  }
  return n;
}

/* NIP_GPU_ENGINE=1|2|3 asks for one engine of the device library (NIPGPU_ENGINE_*: generic
 * join tree, chain, factor by factor); default: the library chooses */
static int configured_engine(void) {
  const char* s = getenv("NIP_GPU_ENGINE");
  const int v = s ? atoi(s) : 0;
  return v >= 1 && v <= 3 ? v : NIPGPU_ENGINE_AUTO;
}

static int parameters_differ(nip_model model, const backend_entry* e) {
  const nipgpu_model_desc* d = e->desc;
  int i, j;
  for (i = 0; i < model->num_of_cliques; i++)
    if (memcmp(e->tables + d->clique_tab_off[i], model->cliques[i]->original_p->data,
               sizeof(double) * (size_t)model->cliques[i]->original_p->size_of_data))
      return 1;
  for (i = 0; i < model->num_of_vars; i++)
    if (model->variables[i]->num_of_parents == 0)
      for (j = 0; j < NIP_CARDINALITY(model->variables[i]); j++)
        if (e->prior[d->var_prior_off[i] + j] != (model->variables[i]->prior ? model->variables[i]->prior[j] : 0.0))
          return 1;
  return 0;
}

static void gather_parameters(nip_model model, const nipgpu_model_desc* d, double* tables, double* prior) {
  int i, j;
  for (i = 0; i < model->num_of_cliques; i++)
    memcpy(tables + d->clique_tab_off[i], model->cliques[i]->original_p->data,
           sizeof(double) * (size_t)model->cliques[i]->original_p->size_of_data);
  for (i = 0; i < model->num_of_vars; i++)
    if (model->variables[i]->num_of_parents == 0)
      for (j = 0; j < NIP_CARDINALITY(model->variables[i]); j++)
        prior[d->var_prior_off[i] + j] = model->variables[i]->prior ? model->variables[i]->prior[j] : 0.0;
}

static void memo_clear(backend_entry* e) {
  int i;
  if (!e->memo) return;
  for (i = 0; i < MEMO_SLOTS; i++) { free(e->memo[i].in); e->memo[i].in = NULL; }
  e->memo_fill = 0;
}

static void destroy_entry(backend_entry* e) {
  int k;
  if (e->group) nipgpu_group_destroy(e->group);
  for (k = 0; k < e->n_dev; k++) nipgpu_model_destroy(e->gm[k]);
  nipgpu_desc_free(e->desc);
  memo_clear(e);
  free(e->memo); free(e->tables); free(e->prior); free(e->seps); free(e->stage);
}

/* Returns the device model of `model`, compiling it on first use.  With `refresh` the host's
 * original_p / priors are compared with what the device holds and pushed again when they
 * changed (make_consistent does not need them: it ships the current tables itself). */
static backend_entry* backend_for(nip_model model, int refresh) {
  backend_entry* e = NULL;
  int i;
  for (i = 0; i < n_backends; i++)
    if (backends[i].model == model) e = &backends[i];
  if (!e) {
    backend_entry fresh;
    const char* memo_env = getenv("NIP_GPU_SLICE_MEMO");
    if (n_backends == cap_backends) {
      int cap = cap_backends ? 2 * cap_backends : 4;
      backend_entry* grown = (backend_entry*)realloc(backends, sizeof(backend_entry) * (size_t)cap);
      if (!grown) { report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY); return NULL; }
      backends = grown;
      cap_backends = cap;
    }
    memset(&fresh, 0, sizeof(fresh));
    fresh.desc = nipgpu_desc_from_model(model);
    if (!fresh.desc) return NULL;
    fresh.model = model;
    fresh.n_dev = 1;
    configured_devices(fresh.dev);
    if (nipgpu_model_create(fresh.desc, fresh.dev[0], configured_engine(), &fresh.gm[0]) != NIPGPU_OK) {
      report_device_error();
      nipgpu_desc_free(fresh.desc);
      return NULL;
    }
    fresh.n_tables = (long)fresh.desc->clique_tab_off[fresh.desc->n_cliques];
    fresh.n_prior = fresh.desc->var_prior_off[fresh.desc->n_vars];
    fresh.tables = (double*)malloc(sizeof(double) * (size_t)(fresh.n_tables > 0 ? fresh.n_tables : 1));
    fresh.prior = (double*)malloc(sizeof(double) * (size_t)(fresh.n_prior > 0 ? fresh.n_prior : 1));
    fresh.seps = nipgpu_model_sepsets(model, &fresh.n_seps);
    if (!fresh.tables || !fresh.prior || !fresh.seps) { destroy_entry(&fresh); return NULL; }
    memcpy(fresh.tables, fresh.desc->clique_tables, sizeof(double) * (size_t)fresh.n_tables);
    memcpy(fresh.prior, fresh.desc->var_prior, sizeof(double) * (size_t)fresh.n_prior);
    for (i = 0; i < fresh.n_seps; i++) fresh.n_msg += fresh.seps[i]->new->size_of_data;
    fresh.stage = (double*)malloc(sizeof(double) * (size_t)(2 * fresh.n_tables + 3 * fresh.n_msg + 1));
    fresh.memo_on = !(memo_env && memo_env[0] == '0') &&
                    2 * fresh.n_tables + 3 * fresh.n_msg <= MEMO_MAX_DOUBLES;
    if (fresh.memo_on) fresh.memo = (memo_slot*)calloc(MEMO_SLOTS, sizeof(memo_slot));
    if (!fresh.stage || (fresh.memo_on && !fresh.memo)) { destroy_entry(&fresh); return NULL; }
    backends[n_backends] = fresh;
    e = &backends[n_backends++];
  }
  if (refresh && parameters_differ(model, e)) {
    int k;
    gather_parameters(model, e->desc, e->tables, e->prior);
    for (k = 0; k < e->n_dev; k++)
      if (nipgpu_model_set_parameters(e->gm[k], e->tables, e->prior) != NIPGPU_OK) {
        report_device_error();
        return NULL;
      }
    e->version++;
  }
  return e;
}

/* Compiles the model on the remaining devices of NIP_GPU_DEVICES and joins them in a group.
 * Returns how many devices EM may use (1 when only one is configured). */
static int ensure_group(backend_entry* e) {
  int dev[NIP_GPU_MAX_DEVICES], n = configured_devices(dev), k;
  if (n <= 1 || e->group) return e->n_dev;
  for (k = 1; k < n; k++) {
    e->dev[k] = dev[k];
    if (nipgpu_model_create(e->desc, dev[k], configured_engine(), &e->gm[k]) != NIPGPU_OK ||
        nipgpu_model_set_parameters(e->gm[k], e->tables, e->prior) != NIPGPU_OK) {
      report_device_error();
      while (k >= 1) { nipgpu_model_destroy(e->gm[k]); e->gm[k--] = NULL; }
      return -1;
    }
  }
  if (nipgpu_group_create(e->gm, n, &e->group) != NIPGPU_OK) {
    report_device_error();
    for (k = 1; k < n; k++) { nipgpu_model_destroy(e->gm[k]); e->gm[k] = NULL; }
    e->group = NULL;
    return -1;
  }
  e->n_dev = n;
  return n;
}

/* Forget the device copy of a model (call before free_model()). */
static void forget_sets_of(nip_model model);
static void drop_parked_of(nip_model model);

void nip_gpu_release(nip_model model) {
  int i;
  forget_sets_of(model);
  for (i = 0; i < n_backends; i++)
    if (backends[i].model == model) {
      if (getenv("NIP_GPU_STATS"))
        fprintf(stderr, "nip gpu backend: make_consistent calls %lu, served from the memo %lu\n",
                backends[i].slice_calls, backends[i].memo_hits);
      destroy_entry(&backends[i]);
      backends[i] = backends[--n_backends];
      return;
    }
}

/* ---- marshalling ---------------------------------------------------------- */
/* NIP_MARK_ON variables only enter evidence (insert_ts_step, src/nip.c:993) */
static unsigned char* marked_mask(nip_model model) {
  unsigned char* m = (unsigned char*)calloc((size_t)(model->num_of_vars > 0 ? model->num_of_vars : 1), 1);
  int i;
  if (m)
    for (i = 0; i < model->num_of_vars; i++) m[i] = (NIP_MARK(model->variables[i]) & NIP_MARK_ON) ? 1 : 0;
  return m;
}

/* Packs the series idx[0..n-1] of `set` (idx == NULL: the first n) and uploads them to device
 * model gm.  *err receives the nip.h error code when NULL is returned. */
static nipgpu_batch* upload_set(nipgpu_model* gm, nip_model model, time_series* set, const int* idx, int n,
                                long* rows_out, int* err) {
  time_series first = n > 0 ? set[idx ? idx[0] : 0] : NULL;
  int n_obs = first ? first->num_of_observed : 0, s, t, k;
  long rows = 0, r = 0;
  int32_t *len, *vars, *data;
  nipgpu_batch* b = NULL;
  *err = NIP_ERROR_OUTOFMEMORY;
  for (s = 0; s < n; s++) {
    time_series ts = set[idx ? idx[s] : s];
    /* every series must carry the same columns in the same order (read_timeseries gives that,
     * src/nip.c:565-612); anything else would enter evidence on the wrong variable */
    if (!ts || ts->model != model || ts->num_of_observed != n_obs) { *err = NIP_ERROR_INVALID_ARGUMENT; return NULL; }
    for (k = 0; k < n_obs; k++)
      if (ts->observed[k] != first->observed[k]) { *err = NIP_ERROR_INVALID_ARGUMENT; return NULL; }
    rows += ts->length;
  }
  len = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  vars = (int32_t*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(int32_t));
  data = (int32_t*)calloc((size_t)(rows * n_obs > 0 ? rows * n_obs : 1), sizeof(int32_t));
  if (len && vars && data) {
    for (k = 0; k < n_obs; k++) vars[k] = nipgpu_var_index(model, first->observed[k]);
    for (s = 0; s < n; s++) {
      time_series ts = set[idx ? idx[s] : s];
      len[s] = ts->length;
      for (t = 0; t < ts->length; t++)
        for (k = 0; k < n_obs; k++) data[r++] = ts->data[t][k];
    }
    if (nipgpu_batch_create(gm, n, len, n_obs, vars, data, &b) != NIPGPU_OK) {
      report_device_error();
      *err = NIP_ERROR_GENERAL;
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
  int i, s, row = 0, rc, err = 0;
  if (!set || n <= 0 || !set[0] || !set[0]->model) return report(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT);
  e = backend_for(set[0]->model, 1);
  if (!e) return report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
  b = upload_set(e->gm[0], e->model, set, NULL, n, &rows, &err);
  if (!b) return report(__FILE__, __LINE__, err);
  mask = marked_mask(e->model);
  q = (int32_t*)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(int32_t));
  for (i = 0; q && i < nvars; i++) {
    q[i] = nipgpu_var_index(e->model, vars[i]);
    row += NIP_CARDINALITY(vars[i]);
  }
  post = (double*)calloc((size_t)(rows * row > 0 ? rows * row : 1), sizeof(double));
  if (!mask || !q || !post) {
    nipgpu_batch_destroy(b); free(mask); free(q); free(post);
    return report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY);
  }
  rc = nipgpu_infer(e->gm[0], b, mask, nvars, q, forward_only, nvars > 0 ? post : NULL, loglik);
  if (rc != NIPGPU_OK) {
    report_device_error();
    report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
  } else
    for (s = 0; s < n; s++) {
      results[s] = new_uncertain_series(vars, nvars, set[s]->length, post + (size_t)r0 * row, row);
      r0 += set[s]->length;
    }
  nipgpu_batch_destroy(b); free(mask); free(q); free(post);
  return rc == NIPGPU_OK ? NIP_NO_ERROR : NIP_ERROR_GENERAL;
}

/* ---- packed data file -> device batch ---------------------
 * A set written by nip_gpu_write_timeseries_bin() (nip_data_bin.c) goes from the file to HBM in
 * one bulk read and one upload: no time_series structs, no per-row allocation, none of
 * read_timeseries()'s two tokenising passes (src/nip.c:512-667).  Results come back flat:
 *   *post     malloc'ed, sum(lengths) rows x (sum of card(vars)) doubles, series after series
 *   *loglik   malloc'ed, one value per series (NULL argument: not wanted)
 *   *lengths  malloc'ed, slices per series;  return value = number of series, 0 on error
 * Every column of the file must name a variable of the model. */
int nip_gpu_smooth_bin(nip_model model, const char* filename, nip_variable vars[], int nvars, int forward_only,
                       double** post, double** loglik, int** lengths) {
  backend_entry* e;
  FILE* f;
  char magic[4];
  int32_t version = 0, n = 0, n_cols = 0, *len = NULL, *col = NULL, *data = NULL, *q = NULL;
  unsigned char* mask = NULL;
  nipgpu_batch* b = NULL;
  double *P = NULL, *L = NULL;
  long rows = 0;
  int k, s, row = 0, ok, rc = NIPGPU_EINVAL;
  if (!model || !filename || !post || !lengths || (nvars > 0 && !vars)) { report(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT); return 0; }
  *post = NULL; *lengths = NULL;
  if (loglik) *loglik = NULL;
  e = backend_for(model, 1);
  if (!e) { report(__FILE__, __LINE__, NIP_ERROR_GENERAL); return 0; }
  f = fopen(filename, "rb");
  if (!f) { report(__FILE__, __LINE__, NIP_ERROR_GENERAL); return 0; }
  ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "NIPB", 4) == 0 && fread(&version, 4, 1, f) == 1 && version == 1 &&
       fread(&n, 4, 1, f) == 1 && fread(&n_cols, 4, 1, f) == 1 && n > 0 && n_cols >= 0;
  if (ok) {
    col = (int32_t*)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(int32_t));
    len = (int32_t*)calloc((size_t)n, sizeof(int32_t));
    ok = col && len;
  }
  for (k = 0; ok && k < n_cols; k++) {   /* column symbols -> model variables */
    int32_t l = 0;
    char sym[4096];
    nip_variable v;
    ok = fread(&l, 4, 1, f) == 1 && l >= 0 && l < (int32_t)sizeof(sym) && fread(sym, 1, (size_t)l, f) == (size_t)l;
    if (!ok) break;
    sym[l] = 0;
    v = model_variable(model, sym);
    ok = v != NULL;
    if (ok) col[k] = nipgpu_var_index(model, v);
  }
  ok = ok && fread(len, 4, (size_t)n, f) == (size_t)n;
  for (s = 0; ok && s < n; s++) { ok = len[s] >= 0; rows += len[s]; }
  if (ok) {
    const size_t cells = (size_t)rows * (size_t)n_cols;
    data = (int32_t*)malloc(sizeof(int32_t) * (cells > 0 ? cells : 1));
    ok = data && fread(data, 4, cells, f) == cells;   /* the whole set in one read */
  }
  fclose(f);
  if (ok) {
    q = (int32_t*)calloc((size_t)(nvars > 0 ? nvars : 1), sizeof(int32_t));
    for (k = 0; q && k < nvars; k++) { q[k] = nipgpu_var_index(model, vars[k]); row += NIP_CARDINALITY(vars[k]); }
    mask = marked_mask(model);
    P = (double*)malloc(sizeof(double) * (size_t)((long)rows * row > 0 ? (long)rows * row : 1));
    L = loglik ? (double*)malloc(sizeof(double) * (size_t)n) : NULL;
    ok = q && mask && P && (!loglik || L);
  }
  if (ok) {
    rc = nipgpu_batch_create(e->gm[0], n, len, n_cols, col, data, &b);
    if (rc == NIPGPU_OK) rc = nipgpu_infer(e->gm[0], b, mask, nvars, q, forward_only, nvars > 0 ? P : NULL, L);
    if (rc != NIPGPU_OK) report_device_error();
    nipgpu_batch_destroy(b);
  }
  free(col); free(data); free(q); free(mask);
  if (!ok || rc != NIPGPU_OK) {
    free(len); free(P); free(L);
    report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
    return 0;
  }
  *post = P; *lengths = len;
  if (loglik) *loglik = L;
  return n;
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

static set_entry* sets = NULL;
static int n_sets = 0, cap_sets = 0;

static void drop_parked(set_entry* se) {
  int i;
  if (se->res)
    for (i = 0; i < se->n; i++)
      if (se->res[i]) { free_uncertainseries(se->res[i]); se->res[i] = NULL; }
  free(se->vars); free(se->mask);
  se->vars = NULL; se->mask = NULL;
  se->valid = 0;
}

/* Tell the backend that set[0..n-1] belong together.  The array and its series must stay alive
 * until nip_gpu_forget_set(set) — call that BEFORE free_timeseries() on any member.  Hosts built
 * with NIP_GPU_WRAP_SETS never call either: their read_timeseries()/free_timeseries() do. */
void nip_gpu_register_set(time_series* set, int n) {
  set_entry* se;
  if (!set || n <= 1) return;
  if (n_sets == cap_sets) {
    int cap = cap_sets ? 2 * cap_sets : 8;
    set_entry* grown = (set_entry*)realloc(sets, sizeof(set_entry) * (size_t)cap);
    if (!grown) { report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY); return; }
    sets = grown;
    cap_sets = cap;
  }
  se = &sets[n_sets];
  memset(se, 0, sizeof(*se));
  se->res = (uncertain_series*)calloc((size_t)n, sizeof(uncertain_series));
  se->ll = (double*)calloc((size_t)n, sizeof(double));
  if (!se->res || !se->ll) { free(se->res); free(se->ll); report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY); return; }
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

#ifdef NIP_GPU_WRAP_SETS
/* a series is going away: the set it belongs to cannot be batched any more */
static void forget_series(time_series ts) {
  int i, k;
  for (i = 0; i < n_sets; i++)
    for (k = 0; k < sets[i].n; k++)
      if (sets[i].set[k] == ts) { nip_gpu_forget_set(sets[i].set); return; }
}
#endif

static uncertain_series batched_or_single(time_series ts, nip_variable vars[], int nvars, int forward_only,
                                          double* loglikelihood) {
  uncertain_series r = NULL;
  set_entry* se = NULL;
  int i, k, idx = -1;
  for (i = 0; i < n_sets && !se; i++)
    for (k = 0; k < sets[i].n; k++)
      if (sets[i].set[k] == ts) { se = &sets[i]; idx = k; break; }
  if (se && ts && ts->model) {
    backend_entry* e = backend_for(ts->model, 1);
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

/* parked posteriors of `model` were computed with parameters that are gone */
static void drop_parked_of(nip_model model) {
  int i;
  for (i = 0; i < n_sets; i++)
    if (sets[i].model == model) drop_parked(&sets[i]);
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
 * variable per slice (it runs unchanged on make_consistent() below).  nip_gpu_generate_set()
 * draws a whole set on the device instead: n ordinary `time_series` in which every model
 * variable is observed (columns in model->variables order), to be freed with
 * free_timeseries().  The random stream is the device library's, not rand().  The set is
 * registered for transparent batching only in a NIP_GPU_WRAP_SETS build (whose
 * free_timeseries() unregisters it); otherwise call nip_gpu_register_set() yourself. */
int nip_gpu_generate_set(nip_model model, int n, int length, unsigned long seed, time_series** results) {
  backend_entry* e;
  int32_t* raw;
  time_series* set;
  int s, t, k, nv, rc;
  if (!model || !results || n <= 0 || length <= 0) {
    report(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT);
    return 0;
  }
  *results = NULL;
  e = backend_for(model, 1);
  if (!e) { report(__FILE__, __LINE__, NIP_ERROR_GENERAL); return 0; }
  nv = model->num_of_vars;
  raw = (int32_t*)calloc((size_t)n * length * nv, sizeof(int32_t));
  set = (time_series*)calloc((size_t)n, sizeof(time_series));
  if (!raw || !set) { free(raw); free(set); report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY); return 0; }
  rc = nipgpu_sample(e->gm[0], n, length, (uint64_t)seed, raw);
  if (rc != NIPGPU_OK) {
    report_device_error();
    report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
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
#ifdef NIP_GPU_WRAP_SETS
  nip_gpu_register_set(set, n);
#endif
  return n;
}

/* ---- EM ---------------------------------------------------------------------
 * Control flow of em_learn (src/nip.c:2076-2250) with E- and M-steps on the
 * device: parameters, expected counts and log-likelihood stay in HBM across
 * iterations; only the scalar log-likelihood and the status come back per
 * iteration, the trained tables once at the end.
 *
 * With NIP_GPU_DEVICES=0,1,...: the series are sharded over the devices (longest first, each to
 * the least loaded device), every device runs the E-step of its shard and ONE ncclAllReduce per
 * iteration sums the expected counts (nipgpu_group_em_estep); the M-step runs on every device. */
static int cmp_desc_len(const void* a, const void* b) {
  const long* x = (const long*)a;
  const long* y = (const long*)b;
  if (x[0] != y[0]) return x[0] > y[0] ? -1 : 1;
  return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

int em_learn(time_series* ts, int n_ts, double threshold, nip_double_list learning_curve) {
  backend_entry* e;
  nipgpu_batch* b[NIP_GPU_MAX_DEVICES];
  nip_model model;
  unsigned char* mask;
  double *init, *tables, *prior;
  double old_ll, ll = -DBL_MAX;
  long n_counts, k = 0;
  int v, i, j, d, ts_steps = 0, status = 0, rc = NIPGPU_OK, iter = 0, n_dev, err = 0;
  if (!ts || n_ts <= 0 || !ts[0] || !ts[0]->model) return report(__FILE__, __LINE__, NIP_ERROR_INVALID_ARGUMENT);
  model = ts[0]->model;
  if (learning_curve != NULL && NIP_LIST_LENGTH(learning_curve) > 0) nip_empty_double_list(learning_curve);
  e = backend_for(model, 1);
  if (!e) return report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
  n_dev = ensure_group(e);
  if (n_dev < 1) return report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
  memset(b, 0, sizeof(b));

  /* shards: longest series first, each to the device with the fewest slices so far */
  {
    long* order = (long*)malloc(sizeof(long) * 2 * (size_t)n_ts);
    int* shard = (int*)malloc(sizeof(int) * (size_t)n_ts * (size_t)n_dev);
    int cnt[NIP_GPU_MAX_DEVICES];
    long load[NIP_GPU_MAX_DEVICES];
    if (!order || !shard) { free(order); free(shard); return report(__FILE__, __LINE__, NIP_ERROR_OUTOFMEMORY); }
    for (i = 0; i < n_ts; i++) { order[2 * i] = ts[i] ? ts[i]->length : 0; order[2 * i + 1] = i; }
    if (n_dev > 1) qsort(order, (size_t)n_ts, 2 * sizeof(long), cmp_desc_len);
    for (d = 0; d < n_dev; d++) { cnt[d] = 0; load[d] = 0; }
    for (i = 0; i < n_ts; i++) {
      int best = 0;
      for (d = 1; d < n_dev; d++)
        if (load[d] < load[best]) best = d;
      shard[(size_t)best * n_ts + cnt[best]++] = (int)order[2 * i + 1];
      load[best] += order[2 * i];
    }
    for (d = 0; d < n_dev && !err; d++) {
      b[d] = upload_set(e->gm[d], model, ts, shard + (size_t)d * n_ts, cnt[d], NULL, &err);
      if (b[d]) err = 0;
    }
    free(order); free(shard);
  }
  mask = marked_mask(model);
  n_counts = (long)nipgpu_model_counts_size(e->gm[0]);
  init = (double*)malloc(sizeof(double) * (size_t)(n_counts > 0 ? n_counts : 1));
  if (err || !mask || !init) {
    for (d = 0; d < n_dev; d++) nipgpu_batch_destroy(b[d]);
    free(mask); free(init);
    return report(__FILE__, __LINE__, err ? err : NIP_ERROR_OUTOFMEMORY);
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

  /* first M-step enters the random parameters */
  rc = e->group ? nipgpu_group_em_mstep(e->group, init) : nipgpu_em_mstep(e->gm[0], init);
  free(init);
  for (;;) {
    if (rc != NIPGPU_OK) break;
    old_ll = ll;
    rc = e->group ? nipgpu_group_em_estep(e->group, b, mask, 1, NULL, &ll, &status)
                  : nipgpu_em_estep(e->gm[0], b[0], mask, 1, NULL, &ll, &status);
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
    /* next M-step straight from the device accumulator */
    rc = e->group ? nipgpu_group_em_mstep(e->group, NULL) : nipgpu_em_mstep(e->gm[0], NULL);
  }
  /* the model keeps the parameters of the last M-step (src/nip.c:2149-2250): bring them home */
  tables = (double*)malloc(sizeof(double) * (size_t)(e->n_tables > 0 ? e->n_tables : 1));
  prior = (double*)malloc(sizeof(double) * (size_t)(e->n_prior > 0 ? e->n_prior : 1));
  if (tables && prior && nipgpu_model_get_parameters(e->gm[0], tables, prior) == NIPGPU_OK) {
    nipgpu_desc_store_parameters(model, e->desc, tables, prior);
    memcpy(e->tables, tables, sizeof(double) * (size_t)e->n_tables);
    memcpy(e->prior, prior, sizeof(double) * (size_t)e->n_prior);
  }
  /* the device parameters changed: nothing computed before this call may be served again */
  e->version++;
  drop_parked_of(model);
  free(tables); free(prior);
  for (d = 0; d < n_dev; d++) nipgpu_batch_destroy(b[d]);
  free(mask);
  if (rc == -NIP_ERROR_BAD_LUCK) return NIP_ERROR_BAD_LUCK;
  if (rc != NIPGPU_OK) {
    report_device_error();
    report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
    if (learning_curve != NULL) nip_empty_double_list(learning_curve);
    return NIP_ERROR_GENERAL;
  }
  return NIP_NO_ERROR;
}

/* ---- single slice -------------------------------------------------------------
 * make_consistent (src/nip.c:1600-1617) on the device, on the tree the host holds: every
 * clique->p and every sepset->new go over in one block, k_jt_propagate runs collect towards
 * cliques[0] and distribute with the reference's message pass (src/nipjointree.c:676-709), and
 * the consistent cliques plus both potentials of every sepset come back, so that the unchanged
 * host code (model_prob_mass, get_probability, nipjoint's reads of clique->p, the next
 * nip_enter_evidence, finish/start_timeslice_message_pass) continues on exactly the state the
 * reference would have left. */
/* hash of the request, eight bytes at a time (the request is an array of doubles) */
static unsigned long long hash_doubles(const double* p, size_t n) {
  unsigned long long h = 0x9e3779b97f4a7c15ULL, w;
  size_t i;
  for (i = 0; i < n; i++) {
    memcpy(&w, p + i, sizeof w);
    h = (h ^ w) * 0xff51afd7ed558ccdULL;
    h ^= h >> 29;
  }
  return h;
}

void make_consistent(nip_model model) {
  backend_entry* e = backend_for(model, 0);
  const nipgpu_model_desc* d;
  double *in, *out;
  size_t n_in, n_out, o;
  unsigned long long h = 0;
  int c, s, slot = -1, hit = 0;
  if (!e) { report(__FILE__, __LINE__, NIP_ERROR_GENERAL); return; }
  d = e->desc;
  n_in = (size_t)(e->n_tables + e->n_msg);
  n_out = (size_t)(e->n_tables + 2 * e->n_msg);
  in = e->stage;
  out = e->stage + n_in;
  for (c = 0; c < model->num_of_cliques; c++)
    memcpy(in + d->clique_tab_off[c], model->cliques[c]->p->data,
           sizeof(double) * (size_t)model->cliques[c]->p->size_of_data);
  for (s = 0, o = (size_t)e->n_tables; s < e->n_seps; s++) {
    memcpy(in + o, e->seps[s]->new->data, sizeof(double) * (size_t)e->seps[s]->new->size_of_data);
    o += (size_t)e->seps[s]->new->size_of_data;
  }
  e->slice_calls++;
  if (e->memo_on) {
    h = hash_doubles(in, n_in);
    for (slot = (int)(h & (MEMO_SLOTS - 1)); e->memo[slot].in; slot = (slot + 1) & (MEMO_SLOTS - 1))
      if (e->memo[slot].hash == h && memcmp(e->memo[slot].in, in, n_in * sizeof(double)) == 0) {
        memcpy(out, e->memo[slot].in + n_in, n_out * sizeof(double));
        hit = 1;
        e->memo_hits++;
        break;
      }
  }
  if (!hit) {
    if (nipgpu_slice_propagate(e->gm[0], in, in + e->n_tables, out, out + e->n_tables,
                               out + e->n_tables + e->n_msg) != NIPGPU_OK) {
      report_device_error();
      report(__FILE__, __LINE__, NIP_ERROR_GENERAL);
      return;
    }
    if (e->memo_on) {
      if (e->memo_fill >= MEMO_MAX_FILL) {   /* full: start over (the working set is tiny or not cacheable) */
        memo_clear(e);
        slot = (int)(h & (MEMO_SLOTS - 1));
      }
      e->memo[slot].in = (double*)malloc((n_in + n_out) * sizeof(double));
      if (e->memo[slot].in) {
        e->memo[slot].hash = h;
        memcpy(e->memo[slot].in, in, n_in * sizeof(double));
        memcpy(e->memo[slot].in + n_in, out, n_out * sizeof(double));
        e->memo_fill++;
      }
    }
  }
  for (c = 0; c < model->num_of_cliques; c++)
    memcpy(model->cliques[c]->p->data, out + d->clique_tab_off[c],
           sizeof(double) * (size_t)model->cliques[c]->p->size_of_data);
  for (s = 0, o = (size_t)e->n_tables; s < e->n_seps; s++) {
    const size_t n = (size_t)e->seps[s]->new->size_of_data;
    memcpy(e->seps[s]->new->data, out + o, sizeof(double) * n);
    memcpy(e->seps[s]->old->data, out + o + (size_t)e->n_msg, sizeof(double) * n);
    o += n;
  }
}
