/* nip_data_bin.c — see nip_data_bin.h */
#include "nip_data_bin.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef NIP_ERROR_GENERAL
#define NIP_ERROR_INVALID_ARGUMENT 3
#define NIP_ERROR_OUTOFMEMORY 4
#define NIP_ERROR_GENERAL 6
#endif

void nip_gpu_register_set(time_series* set, int n);

static int put32(FILE* f, int32_t v) { return fwrite(&v, sizeof v, 1, f) == 1; }
static int get32(FILE* f, int32_t* v) { return fread(v, sizeof *v, 1, f) == 1; }

int nip_gpu_write_timeseries_bin(time_series* set, int n, const char* filename) {
  FILE* f;
  int s, t, k, n_obs, ok = 1;
  if (!set || n <= 0 || !set[0] || !filename) return NIP_ERROR_INVALID_ARGUMENT;
  n_obs = set[0]->num_of_observed;
  for (s = 1; s < n; s++) {
    if (set[s]->num_of_observed != n_obs) return NIP_ERROR_INVALID_ARGUMENT;
    for (k = 0; k < n_obs; k++)
      if (set[s]->observed[k] != set[0]->observed[k]) return NIP_ERROR_INVALID_ARGUMENT;
  }
  f = fopen(filename, "wb");
  if (!f) return NIP_ERROR_GENERAL;
  ok &= fwrite("NIPB", 1, 4, f) == 4;
  ok &= put32(f, 1) && put32(f, n) && put32(f, n_obs);
  for (k = 0; k < n_obs && ok; k++) {
    const char* sym = nip_variable_symbol(set[0]->observed[k]);
    const int32_t len = (int32_t)strlen(sym);
    ok &= put32(f, len) && fwrite(sym, 1, (size_t)len, f) == (size_t)len;
  }
  for (s = 0; s < n && ok; s++) ok &= put32(f, set[s]->length);
  for (s = 0; s < n && ok; s++)
    for (t = 0; t < set[s]->length && ok; t++)
      if (n_obs > 0) ok &= fwrite(set[s]->data[t], sizeof(int), (size_t)n_obs, f) == (size_t)n_obs;
  if (fclose(f) != 0) ok = 0;
  return ok ? NIP_NO_ERROR : NIP_ERROR_GENERAL;
}

static void free_partial(time_series* set, int n) {
  int s;
  if (!set) return;
  for (s = 0; s < n; s++)
    if (set[s]) free_timeseries(set[s]);
  free(set);
}

int nip_gpu_read_timeseries_bin(nip_model model, const char* filename, time_series** results) {
  FILE* f;
  char magic[4];
  int32_t version = 0, n = 0, n_cols = 0, *len = NULL, *row = NULL;
  nip_variable* col_var = NULL;  /* NULL: the model does not know the column */
  time_series* set = NULL;
  int s, t, k, i, h, n_obs = 0, ok = 1;
  if (!model || !filename || !results) return 0;
  *results = NULL;
  f = fopen(filename, "rb");
  if (!f) return 0;
  ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "NIPB", 4) == 0 && get32(f, &version) && version == 1 &&
       get32(f, &n) && get32(f, &n_cols) && n >= 0 && n_cols >= 0;
  if (ok) {
    col_var = (nip_variable*)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(nip_variable));
    len = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    row = (int32_t*)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(int32_t));
    set = (time_series*)calloc((size_t)(n > 0 ? n : 1), sizeof(time_series));
    ok = col_var && len && row && set;
  }
  for (k = 0; k < n_cols && ok; k++) {
    int32_t l = 0;
    char* sym;
    ok = get32(f, &l) && l >= 0 && l < 4096;
    if (!ok) break;
    sym = (char*)calloc((size_t)l + 1, 1);
    ok = sym && fread(sym, 1, (size_t)l, f) == (size_t)l;
    if (ok) {
      col_var[k] = model_variable(model, sym);
      if (col_var[k]) n_obs++;
    }
    free(sym);
  }
  for (s = 0; s < n && ok; s++) ok = get32(f, &len[s]) && len[s] >= 0;
  for (s = 0; s < n && ok; s++) {
    time_series ts = (time_series)calloc(1, sizeof(time_series_struct));
    if (!ts) { ok = 0; break; }
    set[s] = ts;
    ts->model = model;
    ts->length = len[s];
    ts->num_of_observed = n_obs;
    ts->num_of_hidden = model->num_of_vars - n_obs;
    ts->observed = (nip_variable*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(nip_variable));
    ts->hidden = (nip_variable*)calloc((size_t)(ts->num_of_hidden > 0 ? ts->num_of_hidden : 1), sizeof(nip_variable));
    ts->data = (int**)calloc((size_t)(len[s] > 0 ? len[s] : 1), sizeof(int*));
    if (!ts->observed || !ts->hidden || !ts->data) { ok = 0; break; }
    for (k = 0, i = 0; k < n_cols; k++)
      if (col_var[k]) ts->observed[i++] = col_var[k];
    for (i = 0, h = 0; i < model->num_of_vars; i++) {   /* the rest, in model order */
      int seen = 0;
      for (k = 0; k < n_obs; k++)
        if (ts->observed[k] == model->variables[i]) seen = 1;
      if (!seen && h < ts->num_of_hidden) ts->hidden[h++] = model->variables[i];
    }
    for (t = 0; t < len[s] && ok; t++) {
      ts->data[t] = (int*)calloc((size_t)(n_obs > 0 ? n_obs : 1), sizeof(int));
      ok = ts->data[t] && (n_cols == 0 || fread(row, sizeof(int32_t), (size_t)n_cols, f) == (size_t)n_cols);
      for (k = 0, i = 0; k < n_cols && ok; k++)
        if (col_var[k]) {
          ok = row[k] >= -1 && row[k] < NIP_CARDINALITY(col_var[k]);
          ts->data[t][i++] = row[k];
        }
    }
  }
  fclose(f);
  free(col_var); free(len); free(row);
  if (!ok) {
    nip_report_error(__FILE__, __LINE__, NIP_ERROR_GENERAL, 1);
    free_partial(set, n);
    return 0;
  }
  *results = set;
#ifdef NIP_GPU_WRAP_SETS   /* only where free_timeseries() is the wrapper that unregisters the set */
  nip_gpu_register_set(set, n);
#endif
  return n;
}
