/* nip_model_export.c — see nip_model_export.h.
 *
 * Field-by-field provenance (reference file:line):
 *   var_card / var_flags / parents   src/nipvariable.h:51-78
 *   var_family                       first clique, in array order, that holds
 *                                    the variable and all its parents
 *                                    (nip_find_family → nip_find_clique,
 *                                    src/nipjointree.c:967-1064)
 *   clique_vars / tables             nip_clique_struct, src/nipjointree.h:43-51
 *   sepsets                          nip_sepset_struct, src/nipjointree.h:55-63;
 *                                    each sepset hangs in both neighbours'
 *                                    linked lists (nip_confirm_sepset,
 *                                    src/nipjointree.c:176-199)
 *   interface arrays, in/out clique  src/nip.h:88-99, src/nip.c:205-264
 */
#include "nip_model_export.h"

#include <stdlib.h>
#include <string.h>

int nipgpu_var_index(nip_model model, nip_variable v) {
  int i;
  for (i = 0; i < model->num_of_vars; i++)
    if (model->variables[i] == v) return i;
  return -1;
}

static int clique_index(nip_model model, nip_clique c) {
  int i;
  for (i = 0; i < model->num_of_cliques; i++)
    if (model->cliques[i] == c) return i;
  return -1;
}

static int clique_holds(nip_clique c, nip_variable v) {
  int k;
  for (k = 0; k < NIP_DIMENSIONALITY(c->p); k++)
    if (c->variables[k] == v) return 1;
  return 0;
}

static int family_clique_of(nip_model model, nip_variable v) {
  int c, j, ok;
  for (c = 0; c < model->num_of_cliques; c++) {
    ok = clique_holds(model->cliques[c], v);
    for (j = 0; ok && j < v->num_of_parents; j++)
      ok = clique_holds(model->cliques[c], v->parents[j]);
    if (ok) return c;
  }
  return -1;
}

nip_sepset* nipgpu_model_sepsets(nip_model model, int* n) {
  nip_sepset* seen;
  nip_sepset_link l;
  int i, j, ns = 0, cap = 0;
  for (i = 0; i < model->num_of_cliques; i++)
    for (l = model->cliques[i]->sepsets; l; l = l->fwd) cap++;
  seen = (nip_sepset*)calloc((size_t)(cap > 0 ? cap : 1), sizeof(nip_sepset));
  if (!seen) return NULL;
  for (i = 0; i < model->num_of_cliques; i++)
    for (l = model->cliques[i]->sepsets; l; l = l->fwd) {
      for (j = 0; j < ns; j++)
        if (seen[j] == (nip_sepset)l->data) break;
      if (j == ns) seen[ns++] = (nip_sepset)l->data;
    }
  *n = ns;
  return seen;
}

#define ALLOC(ptr, n, type)                                   \
  do {                                                        \
    (ptr) = (type*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(type)); \
    if (!(ptr)) goto fail;                                    \
  } while (0)

nipgpu_model_desc* nipgpu_desc_from_model(nip_model model) {
  nipgpu_model_desc* d;
  int32_t *card = NULL, *flags = NULL, *poff = NULL, *parents = NULL, *family = NULL;
  int32_t *prior_off = NULL, *cvoff = NULL, *cvars = NULL, *scl = NULL, *svoff = NULL;
  int32_t *svars = NULL, *adjoff = NULL, *adj = NULL, *outg = NULL, *prev = NULL;
  int64_t* toff = NULL;
  double *prior = NULL, *tables = NULL;
  nip_sepset* seen = NULL;
  int nv, nc, i, j, k, ns = 0, n, total;
  int64_t tab_total;
  nip_sepset_link l;

  if (!model) return NULL;
  nv = model->num_of_vars;
  nc = model->num_of_cliques;
  d = (nipgpu_model_desc*)calloc(1, sizeof(*d));
  if (!d) return NULL;

  /* variables */
  ALLOC(card, nv, int32_t); ALLOC(flags, nv, int32_t); ALLOC(family, nv, int32_t);
  ALLOC(poff, nv + 1, int32_t); ALLOC(prior_off, nv + 1, int32_t);
  total = 0; n = 0;
  for (i = 0; i < nv; i++) {
    nip_variable v = model->variables[i];
    card[i] = NIP_CARDINALITY(v);
    flags[i] = v->interface_status;
    poff[i] = total; total += v->num_of_parents;
    prior_off[i] = n; if (v->num_of_parents == 0) n += NIP_CARDINALITY(v);
    family[i] = family_clique_of(model, v);
    if (family[i] < 0) goto fail;
  }
  poff[nv] = total; prior_off[nv] = n;
  ALLOC(parents, total, int32_t); ALLOC(prior, n, double);
  for (i = 0; i < nv; i++) {
    nip_variable v = model->variables[i];
    for (j = 0; j < v->num_of_parents; j++) {
      parents[poff[i] + j] = nipgpu_var_index(model, v->parents[j]);
      if (parents[poff[i] + j] < 0) goto fail;
    }
    if (v->num_of_parents == 0)
      for (j = 0; j < NIP_CARDINALITY(v); j++)
        prior[prior_off[i] + j] = v->prior ? v->prior[j] : 0.0;
  }

  /* cliques */
  ALLOC(cvoff, nc + 1, int32_t); ALLOC(toff, nc + 1, int64_t); ALLOC(adjoff, nc + 1, int32_t);
  total = 0; tab_total = 0; n = 0;
  for (i = 0; i < nc; i++) {
    nip_clique c = model->cliques[i];
    cvoff[i] = total; total += NIP_DIMENSIONALITY(c->p);
    toff[i] = tab_total; tab_total += c->original_p->size_of_data;
    adjoff[i] = n;
    for (l = c->sepsets; l; l = l->fwd) n++;
  }
  cvoff[nc] = total; toff[nc] = tab_total; adjoff[nc] = n;
  ALLOC(cvars, total, int32_t); ALLOC(tables, tab_total, double); ALLOC(adj, n, int32_t);
  seen = (nip_sepset*)calloc((size_t)(n > 0 ? n : 1), sizeof(nip_sepset));
  if (!seen) goto fail;
  for (i = 0; i < nc; i++) {
    nip_clique c = model->cliques[i];
    for (k = 0; k < NIP_DIMENSIONALITY(c->p); k++)
      cvars[cvoff[i] + k] = nipgpu_var_index(model, c->variables[k]);
    memcpy(tables + toff[i], c->original_p->data,
           sizeof(double) * (size_t)c->original_p->size_of_data);
    for (l = c->sepsets, k = 0; l; l = l->fwd, k++) {
      nip_sepset s = (nip_sepset)l->data;
      for (j = 0; j < ns; j++)
        if (seen[j] == s) break;
      if (j == ns) seen[ns++] = s;
      adj[adjoff[i] + k] = j;
    }
  }

  /* sepsets */
  ALLOC(scl, 2 * ns, int32_t); ALLOC(svoff, ns + 1, int32_t);
  total = 0;
  for (j = 0; j < ns; j++) {
    scl[2 * j] = clique_index(model, seen[j]->first_neighbour);
    scl[2 * j + 1] = clique_index(model, seen[j]->second_neighbour);
    if (scl[2 * j] < 0 || scl[2 * j + 1] < 0) goto fail;
    svoff[j] = total; total += NIP_DIMENSIONALITY(seen[j]->old);
  }
  svoff[ns] = total;
  ALLOC(svars, total, int32_t);
  for (j = 0; j < ns; j++)
    for (k = 0; k < NIP_DIMENSIONALITY(seen[j]->old); k++)
      svars[svoff[j] + k] = nipgpu_var_index(model, seen[j]->variables[k]);

  /* interface */
  n = model->outgoing_interface_size;
  ALLOC(outg, n, int32_t); ALLOC(prev, n, int32_t);
  for (i = 0; i < n; i++) {
    outg[i] = nipgpu_var_index(model, model->outgoing_interface[i]);
    prev[i] = nipgpu_var_index(model, model->previous_outgoing_interface[i]);
  }

  d->n_vars = nv; d->var_card = card; d->var_flags = flags; d->var_parent_off = poff;
  d->var_parents = parents; d->var_family = family; d->var_prior_off = prior_off;
  d->var_prior = prior;
  d->n_cliques = nc; d->clique_var_off = cvoff; d->clique_vars = cvars;
  d->clique_tab_off = toff; d->clique_tables = tables;
  d->n_sepsets = ns; d->sepset_cliques = scl; d->sepset_var_off = svoff;
  d->sepset_vars = svars; d->clique_adj_off = adjoff; d->clique_adj = adj;
  d->n_interface = n; d->outgoing = outg; d->prev_outgoing = prev;
  d->in_clique = n > 0 ? clique_index(model, model->in_clique) : -1;
  d->out_clique = n > 0 ? clique_index(model, model->out_clique) : -1;
  free(seen);
  return d;

fail:
  free(card); free(flags); free(poff); free(parents); free(family); free(prior_off);
  free(prior); free(cvoff); free(cvars); free(toff); free(tables); free(scl); free(svoff);
  free(svars); free(adjoff); free(adj); free(outg); free(prev); free(seen); free(d);
  return NULL;
}

void nipgpu_desc_free(nipgpu_model_desc* d) {
  if (!d) return;
  free((void*)d->var_card); free((void*)d->var_flags); free((void*)d->var_parent_off);
  free((void*)d->var_parents); free((void*)d->var_family); free((void*)d->var_prior_off);
  free((void*)d->var_prior); free((void*)d->clique_var_off); free((void*)d->clique_vars);
  free((void*)d->clique_tab_off); free((void*)d->clique_tables);
  free((void*)d->sepset_cliques); free((void*)d->sepset_var_off); free((void*)d->sepset_vars);
  free((void*)d->clique_adj_off); free((void*)d->clique_adj);
  free((void*)d->outgoing); free((void*)d->prev_outgoing);
  free(d);
}

void nipgpu_desc_store_parameters(nip_model model, const nipgpu_model_desc* d,
                                  const double* clique_tables, const double* var_prior) {
  int i, j;
  for (i = 0; i < model->num_of_cliques; i++) {
    nip_clique c = model->cliques[i];
    size_t bytes = sizeof(double) * (size_t)c->original_p->size_of_data;
    memcpy(c->original_p->data, clique_tables + d->clique_tab_off[i], bytes);
    memcpy(c->p->data, clique_tables + d->clique_tab_off[i], bytes);
  }
  for (i = 0; i < model->num_of_vars; i++) {
    nip_variable v = model->variables[i];
    if (v->num_of_parents == 0 && v->prior)
      for (j = 0; j < NIP_CARDINALITY(v); j++)
        v->prior[j] = var_prior[d->var_prior_off[i] + j];
  }
}
