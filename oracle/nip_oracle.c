/* TEST INFRASTRUCTURE — CPU oracle (see nip_oracle.h).  Not product code.
 *
 * Restates, on the flat nipgpu_model_desc, the algorithm of
 *   src/nippotential.c  (table algebra),
 *   src/nipjointree.c   (message passing, evidence, mass) and
 *   src/nip.c           (slice loops, EM)
 * of manuelschmidt/nip.  Every function names the reference lines it follows.
 * Summation and multiply/divide ORDER is kept identical to the reference
 * (ascending flat source index; multiply before divide) so that the oracle is
 * bit-comparable with oracle/_ref wherever the same tables go in.
 */
#include "nip_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 32
#define MARK_OFF_BIT 1 /* NIP_MARK_OFF, src/nipvariable.h:37 */
#define MARK_ON_BIT 2  /* NIP_MARK_ON,  src/nipvariable.h:38 */

static int table_size(int ndim, const int* card) {
  int i, n = 1;
  for (i = 0; i < ndim; i++) n *= card[i];
  return n;
}

/* Walks the source table in flat order (dimension 0 fastest,
 * src/nippotential.c:58-68) while tracking the flat index of the projected
 * entry in a sub-table whose k-th dimension is source dimension mapping[k].
 * Replaces nip_inverse_mapping + nip_choose_potential_indices +
 * nip_get_potential_pointer (src/nippotential.c:251-264, 72-81, 58-68). */
typedef struct {
  int ndim, idx[MAXD], card[MAXD], step[MAXD]; /* step = sub-table stride of that dim (0 if dropped) */
  int sub;                                      /* current flat index in the sub-table */
} walker;

static void walker_init(walker* w, int ndim, const int* card, int sub_ndim, const int* sub_card,
                        const int* mapping) {
  int k, stride = 1;
  w->ndim = ndim;
  w->sub = 0;
  for (k = 0; k < ndim; k++) { w->idx[k] = 0; w->card[k] = card[k]; w->step[k] = 0; }
  for (k = 0; k < sub_ndim; k++) { w->step[mapping[k]] = stride; stride *= sub_card[k]; }
}

static void walker_next(walker* w) {
  int k;
  for (k = 0; k < w->ndim; k++) {
    w->idx[k]++;
    w->sub += w->step[k];
    if (w->idx[k] < w->card[k]) return;
    w->sub -= w->step[k] * w->card[k];
    w->idx[k] = 0;
  }
}

/* nip_general_marginalise, src/nippotential.c:267-311 */
int orc_general_marginalise(const double* src, int src_ndim, const int* src_card, double* dst,
                            int dst_ndim, const int* dst_card, const int* mapping) {
  int n = table_size(src_ndim, src_card), m, i;
  walker w;
  if (dst_ndim > src_ndim) return 1;
  if (dst_ndim == 0) { /* :276-282 scalar destination = total sum */
    dst[0] = 0;
    for (i = 0; i < n; i++) dst[0] += src[i];
    return 0;
  }
  m = table_size(dst_ndim, dst_card);
  for (i = 0; i < m; i++) dst[i] = 0.0; /* :285 */
  walker_init(&w, src_ndim, src_card, dst_ndim, dst_card, mapping);
  for (i = 0; i < n; i++) { /* :288-303 */
    dst[w.sub] += src[i];
    walker_next(&w);
  }
  return 0;
}

/* nip_total_marginalise, src/nippotential.c:314-346 */
int orc_total_marginalise(const double* src, int ndim, const int* card, double* dst, int variable) {
  int n = table_size(ndim, card), i, stride = 1, k;
  if (ndim == 0) { dst[0] = src[0]; return 0; }
  if (variable < 0 || variable >= ndim) return 1;
  for (k = 0; k < variable; k++) stride *= card[k];
  for (i = 0; i < card[variable]; i++) dst[i] = 0.0;
  for (i = 0; i < n; i++) dst[(i / stride) % card[variable]] += src[i];
  return 0;
}

/* nip_update_potential, src/nippotential.c:436-496.  num or den may be NULL.
 * Zero denominator ⇒ the entry becomes 0 (:486-491). */
int orc_update_potential(const double* num, const double* den, int sub_ndim, const int* sub_card,
                         double* target, int ndim, const int* card, const int* mapping) {
  int n = table_size(ndim, card), i;
  walker w;
  if (!num && !den) return 1;
  if (sub_ndim == 0) { /* :459-472 */
    for (i = 0; i < n; i++) {
      if (num) target[i] *= num[0];
      if (den) { if (den[0]) target[i] /= den[0]; else target[i] = 0; }
    }
    return 0;
  }
  walker_init(&w, ndim, card, sub_ndim, sub_card, mapping);
  for (i = 0; i < n; i++) { /* :475-493 */
    if (num) target[i] *= num[w.sub];
    if (den) { if (den[w.sub] != 0) target[i] /= den[w.sub]; else target[i] = 0; }
    walker_next(&w);
  }
  return 0;
}

/* nip_update_evidence, src/nippotential.c:499-522.  Zero denominator ⇒ the
 * division is SKIPPED (unlike update_potential). */
int orc_update_evidence(const double* num, const double* den, double* target, int ndim,
                        const int* card, int var) {
  int n = table_size(ndim, card), i, k, stride = 1, s;
  for (k = 0; k < var; k++) stride *= card[k];
  for (i = 0; i < n; i++) {
    s = (i / stride) % card[var];
    target[i] *= num[s];
    if (den && den[s] != 0) target[i] /= den[s];
  }
  return 0;
}

/* nip_init_potential, src/nippotential.c:525-564 */
int orc_init_potential(const double* probs, int sub_ndim, const int* sub_card, double* target,
                       int ndim, const int* card, const int* mapping) {
  int n = table_size(ndim, card), i;
  walker w;
  if (!mapping) {
    if (table_size(sub_ndim, sub_card) != n) return 1;
    for (i = 0; i < n; i++) target[i] *= probs[i];
    return 0;
  }
  if (sub_ndim == 0) return 0;
  walker_init(&w, ndim, card, sub_ndim, sub_card, mapping);
  for (i = 0; i < n; i++) { target[i] *= probs[w.sub]; walker_next(&w); }
  return 0;
}

/* nip_normalise_array, src/nippotential.c:349-360: zero sum ⇒ untouched */
void orc_normalise_array(double* a, int n) {
  double sum = 0;
  int i;
  for (i = 0; i < n; i++) sum += a[i];
  if (sum == 0) return;
  for (i = 0; i < n; i++) a[i] /= sum;
}

/* nip_normalise_cpd, src/nippotential.c:373-383: blocks of card[0] */
void orc_normalise_cpd(double* a, int size, int card0) {
  int i;
  for (i = 0; i < size; i += card0) orc_normalise_array(a + i, card0);
}

/* ======================================================================= */
struct orc_model {
  int nv, nc, ns, nif, in_clique, out_clique;
  int *card, *flags, *poff, *parents, *family, *prior_off;
  double* prior;
  double** lik;
  int* prior_entered;
  int *cvoff, *cvars, *csize;
  long* toff;
  double *orig, *p;
  int *scl, *svoff, *svars, *ssize, *adjoff, *adj;
  double **s_old, **s_new;
  int *outg, *prev;
  int* cmark;
  long* coff; /* per-variable offset of the family count table */
};

static int* dup_i(const int32_t* a, long n) {
  int* r = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  long i;
  for (i = 0; i < n; i++) r[i] = a[i];
  return r;
}

static int cdim(const orc_model* m, int c) { return m->cvoff[c + 1] - m->cvoff[c]; }
static const int* cvars(const orc_model* m, int c) { return m->cvars + m->cvoff[c]; }
static int sdim(const orc_model* m, int s) { return m->svoff[s + 1] - m->svoff[s]; }
static const int* svars(const orc_model* m, int s) { return m->svars + m->svoff[s]; }

static void cards_of(const orc_model* m, const int* vars, int n, int* out) {
  int i;
  for (i = 0; i < n; i++) out[i] = m->card[vars[i]];
}

/* nip_mapper, src/nipvariable.c:560-589: position of each subset variable
 * inside `set` */
static void mapper(const int* set, int nset, const int* subset, int nsub, int* mapping) {
  int i, j;
  for (i = 0; i < nsub; i++)
    for (j = 0; j < nset; j++)
      if (subset[i] == set[j]) { mapping[i] = j; break; }
}

static int var_pos_in_clique(const orc_model* m, int c, int var) {
  int k;
  for (k = 0; k < cdim(m, c); k++)
    if (cvars(m, c)[k] == var) return k;
  return -1;
}

orc_model* orc_model_new(const nipgpu_model_desc* d) {
  orc_model* m = (orc_model*)calloc(1, sizeof(*m));
  int i, c, s;
  long n;
  m->nv = d->n_vars; m->nc = d->n_cliques; m->ns = d->n_sepsets; m->nif = d->n_interface;
  m->in_clique = d->in_clique; m->out_clique = d->out_clique;
  m->card = dup_i(d->var_card, m->nv); m->flags = dup_i(d->var_flags, m->nv);
  m->poff = dup_i(d->var_parent_off, m->nv + 1);
  m->parents = dup_i(d->var_parents, m->poff[m->nv]);
  m->family = dup_i(d->var_family, m->nv);
  m->prior_off = dup_i(d->var_prior_off, m->nv + 1);
  n = m->prior_off[m->nv];
  m->prior = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  memcpy(m->prior, d->var_prior, sizeof(double) * (size_t)n);
  m->lik = (double**)calloc((size_t)m->nv, sizeof(double*));
  m->prior_entered = (int*)calloc((size_t)m->nv, sizeof(int));
  for (i = 0; i < m->nv; i++) m->lik[i] = (double*)calloc((size_t)m->card[i], sizeof(double));
  m->cvoff = dup_i(d->clique_var_off, m->nc + 1);
  m->cvars = dup_i(d->clique_vars, m->cvoff[m->nc]);
  m->toff = (long*)calloc((size_t)m->nc + 1, sizeof(long));
  m->csize = (int*)calloc((size_t)m->nc, sizeof(int));
  for (c = 0; c <= m->nc; c++) m->toff[c] = (long)d->clique_tab_off[c];
  for (c = 0; c < m->nc; c++) m->csize[c] = (int)(m->toff[c + 1] - m->toff[c]);
  n = m->toff[m->nc];
  m->orig = (double*)calloc((size_t)n, sizeof(double));
  m->p = (double*)calloc((size_t)n, sizeof(double));
  memcpy(m->orig, d->clique_tables, sizeof(double) * (size_t)n);
  m->scl = dup_i(d->sepset_cliques, 2L * m->ns);
  m->svoff = dup_i(d->sepset_var_off, m->ns + 1);
  m->svars = dup_i(d->sepset_vars, m->svoff[m->ns]);
  m->ssize = (int*)calloc((size_t)(m->ns > 0 ? m->ns : 1), sizeof(int));
  m->s_old = (double**)calloc((size_t)(m->ns > 0 ? m->ns : 1), sizeof(double*));
  m->s_new = (double**)calloc((size_t)(m->ns > 0 ? m->ns : 1), sizeof(double*));
  for (s = 0; s < m->ns; s++) {
    int sz = 1;
    for (i = 0; i < sdim(m, s); i++) sz *= m->card[svars(m, s)[i]];
    m->ssize[s] = sz;
    m->s_old[s] = (double*)calloc((size_t)sz, sizeof(double));
    m->s_new[s] = (double*)calloc((size_t)sz, sizeof(double));
  }
  m->adjoff = dup_i(d->clique_adj_off, m->nc + 1);
  m->adj = dup_i(d->clique_adj, m->adjoff[m->nc]);
  m->outg = dup_i(d->outgoing, m->nif);
  m->prev = dup_i(d->prev_outgoing, m->nif);
  m->cmark = (int*)calloc((size_t)m->nc, sizeof(int));
  m->coff = (long*)calloc((size_t)m->nv + 1, sizeof(long));
  for (i = 0; i < m->nv; i++) {
    long sz = m->card[i];
    int j;
    for (j = m->poff[i]; j < m->poff[i + 1]; j++) sz *= m->card[m->parents[j]];
    m->coff[i + 1] = m->coff[i] + sz;
  }
  orc_reset_model(m);
  return m;
}

void orc_model_free(orc_model* m) {
  int i;
  if (!m) return;
  for (i = 0; i < m->nv; i++) free(m->lik[i]);
  for (i = 0; i < m->ns; i++) { free(m->s_old[i]); free(m->s_new[i]); }
  free(m->card); free(m->flags); free(m->poff); free(m->parents); free(m->family);
  free(m->prior_off); free(m->prior); free(m->lik); free(m->prior_entered); free(m->cvoff);
  free(m->cvars); free(m->csize); free(m->toff); free(m->orig); free(m->p); free(m->scl);
  free(m->svoff); free(m->svars); free(m->ssize); free(m->adjoff); free(m->adj);
  free(m->s_old); free(m->s_new); free(m->outg); free(m->prev); free(m->cmark); free(m->coff);
  free(m);
}

long orc_counts_size(const orc_model* m) { return m->coff[m->nv]; }

/* ---- join tree -------------------------------------------------------- */
/* nip_message_pass, src/nipjointree.c:676-709 */
static void message_pass(orc_model* m, int c1, int s, int c2) {
  int map[MAXD], ccard[MAXD], scard[MAXD];
  double* t = m->s_old[s];
  m->s_old[s] = m->s_new[s]; /* :682-685 swap */
  m->s_new[s] = t;
  cards_of(m, svars(m, s), sdim(m, s), scard);
  cards_of(m, cvars(m, c1), cdim(m, c1), ccard);
  mapper(cvars(m, c1), cdim(m, c1), svars(m, s), sdim(m, s), map);
  orc_general_marginalise(m->p + m->toff[c1], cdim(m, c1), ccard, m->s_new[s], sdim(m, s), scard,
                          map); /* :690-694 */
  cards_of(m, cvars(m, c2), cdim(m, c2), ccard);
  mapper(cvars(m, c2), cdim(m, c2), svars(m, s), sdim(m, s), map);
  orc_update_potential(m->s_new[s], m->s_old[s], sdim(m, s), scard, m->p + m->toff[c2],
                       cdim(m, c2), ccard, map); /* :701-705 */
}

/* nip_collect_evidence, src/nipjointree.c:630-673 (c1 < 0 = no parent) */
static void collect(orc_model* m, int c1, int s12, int c2) {
  int l, s;
  m->cmark[c2] = 1;
  for (l = m->adjoff[c2]; l < m->adjoff[c2 + 1]; l++) {
    s = m->adj[l];
    if (!m->cmark[m->scl[2 * s]]) collect(m, c2, s, m->scl[2 * s]);
    if (!m->cmark[m->scl[2 * s + 1]]) collect(m, c2, s, m->scl[2 * s + 1]);
  }
  if (c1 >= 0) message_pass(m, c2, s12, c1);
}

/* nip_distribute_evidence, src/nipjointree.c:580-627 */
static void distribute(orc_model* m, int c) {
  int l, s;
  m->cmark[c] = 1;
  for (l = m->adjoff[c]; l < m->adjoff[c + 1]; l++) {
    s = m->adj[l];
    if (!m->cmark[m->scl[2 * s]]) message_pass(m, c, s, m->scl[2 * s]);
    else if (!m->cmark[m->scl[2 * s + 1]]) message_pass(m, c, s, m->scl[2 * s + 1]);
  }
  for (l = m->adjoff[c]; l < m->adjoff[c + 1]; l++) {
    s = m->adj[l];
    if (!m->cmark[m->scl[2 * s]]) distribute(m, m->scl[2 * s]);
    else if (!m->cmark[m->scl[2 * s + 1]]) distribute(m, m->scl[2 * s + 1]);
  }
}

static void unmark_all(orc_model* m) { memset(m->cmark, 0, sizeof(int) * (size_t)m->nc); }

/* make_consistent, src/nip.c:1600-1617 */
void orc_make_consistent(orc_model* m) {
  unmark_all(m);
  collect(m, -1, -1, 0);
  unmark_all(m);
  distribute(m, 0);
}

/* nip_join_tree_dfs with retract callbacks, src/nipjointree.c:1108-1153, 1089-1105 */
static void retract_dfs(orc_model* m, int c) {
  int l, s, k, nb;
  m->cmark[c] = 1;
  memcpy(m->p + m->toff[c], m->orig + m->toff[c], sizeof(double) * (size_t)m->csize[c]);
  for (l = m->adjoff[c]; l < m->adjoff[c + 1]; l++) {
    s = m->adj[l];
    nb = !m->cmark[m->scl[2 * s]] ? m->scl[2 * s] : (!m->cmark[m->scl[2 * s + 1]] ? m->scl[2 * s + 1] : -1);
    if (nb < 0) continue;
    for (k = 0; k < m->ssize[s]; k++) { m->s_old[s][k] = 1; m->s_new[s][k] = 1; }
    retract_dfs(m, nb);
  }
}

/* nip_global_retraction, src/nipjointree.c:791-817 */
static void global_retraction(orc_model* m) {
  int v, c, ccard[MAXD];
  unmark_all(m);
  retract_dfs(m, 0);
  for (v = 0; v < m->nv; v++) { /* re-enter every likelihood (:805-814) */
    c = m->family[v];
    cards_of(m, cvars(m, c), cdim(m, c), ccard);
    orc_update_evidence(m->lik[v], NULL, m->p + m->toff[c], cdim(m, c), ccard,
                        var_pos_in_clique(m, c, v));
  }
}

/* reset_model, src/nip.c:61-73 */
void orc_reset_model(orc_model* m) {
  int v, i;
  for (v = 0; v < m->nv; v++) {
    for (i = 0; i < m->card[v]; i++) m->lik[v][i] = 1;
    m->prior_entered[v] = 0;
  }
  global_retraction(m);
}

/* total_reset, src/nip.c:76-85 */
void orc_total_reset(orc_model* m) {
  long i;
  for (i = 0; i < m->toff[m->nc]; i++) m->orig[i] = 1.0;
  orc_reset_model(m);
}

/* nip_enter_prior, src/nipjointree.c:904-943: zero vector ⇒ not entered */
static int enter_prior(orc_model* m, int v) {
  const double* pr = m->prior + m->prior_off[v];
  int c = m->family[v], i, zero = 1, ccard[MAXD];
  for (i = 0; i < m->card[v]; i++)
    if (pr[i] > 0) zero = 0;
  if (zero) return 1;
  cards_of(m, cvars(m, c), cdim(m, c), ccard);
  orc_update_evidence(pr, NULL, m->p + m->toff[c], cdim(m, c), ccard, var_pos_in_clique(m, c, v));
  return 0;
}

/* use_priors, src/nip.c:88-119: parentless variables in variable order; with
 * history the OLD_OUTGOING ones are skipped */
void orc_use_priors(orc_model* m, int has_history) {
  int v;
  for (v = 0; v < m->nv; v++) {
    if (m->poff[v + 1] != m->poff[v]) continue; /* has parents */
    if (m->prior_entered[v]) continue;
    if (!has_history || !(m->flags[v] & NIPGPU_IF_OLD_OUTGOING)) {
      enter_prior(m, v);
      m->prior_entered[v] = 1;
    }
  }
}

/* nip_enter_evidence, src/nipjointree.c:859-901 */
int orc_enter_evidence(orc_model* m, int v, const double* ev) {
  int c = m->family[v], i, retraction = 0, ccard[MAXD];
  for (i = 0; i < m->card[v]; i++)
    if (m->lik[v][i] == 0 && ev[i] != 0) retraction = 1;
  if (!retraction) {
    cards_of(m, cvars(m, c), cdim(m, c), ccard);
    orc_update_evidence(ev, m->lik[v], m->p + m->toff[c], cdim(m, c), ccard,
                        var_pos_in_clique(m, c, v));
  }
  for (i = 0; i < m->card[v]; i++) m->lik[v][i] = ev[i];
  if (retraction) global_retraction(m);
  return 0;
}

/* nip_enter_index_observation, src/nipjointree.c:832-856 */
int orc_enter_index_observation(orc_model* m, int v, int index) {
  double* e;
  int i, r;
  if (index < 0) return 0;
  e = (double*)calloc((size_t)m->card[v], sizeof(double));
  for (i = 0; i < m->card[v]; i++) e[i] = (i == index) ? 1 : 0;
  r = orc_enter_evidence(m, v, e);
  free(e);
  return r;
}

static void mass_dfs(orc_model* m, int c, double* acc) {
  int l, s, k, nb;
  double x = 0;
  m->cmark[c] = 1;
  for (k = 0; k < m->csize[c]; k++) x += m->p[m->toff[c] + k];
  *acc += x;
  for (l = m->adjoff[c]; l < m->adjoff[c + 1]; l++) {
    s = m->adj[l];
    nb = !m->cmark[m->scl[2 * s]] ? m->scl[2 * s] : (!m->cmark[m->scl[2 * s + 1]] ? m->scl[2 * s + 1] : -1);
    if (nb < 0) continue;
    x = 0;
    for (k = 0; k < m->ssize[s]; k++) x += m->s_new[s][k];
    *acc -= x;
    mass_dfs(m, nb, acc);
  }
}

/* nip_probability_mass, src/nipjointree.c:1156-1188: Σ cliques − Σ sepsets(new) */
double orc_prob_mass(orc_model* m) {
  double r = 0;
  unmark_all(m);
  mass_dfs(m, 0, &r);
  return r;
}

/* get_probability, src/nip.c:2261-2298 (= nip_marginalise_clique + normalise) */
int orc_marginal(orc_model* m, int v, double* out) {
  int c = m->family[v], ccard[MAXD];
  cards_of(m, cvars(m, c), cdim(m, c), ccard);
  orc_total_marginalise(m->p + m->toff[c], cdim(m, c), ccard, out, var_pos_in_clique(m, c, v));
  orc_normalise_array(out, m->card[v]);
  return 0;
}

void orc_get_clique(orc_model* m, int c, int original, double* out) {
  memcpy(out, (original ? m->orig : m->p) + m->toff[c], sizeof(double) * (size_t)m->csize[c]);
}

void orc_get_parameters(orc_model* m, double* tables, double* prior) {
  memcpy(tables, m->orig, sizeof(double) * (size_t)m->toff[m->nc]);
  memcpy(prior, m->prior, sizeof(double) * (size_t)m->prior_off[m->nv]);
}

void orc_set_parameters(orc_model* m, const double* tables, const double* prior) {
  memcpy(m->orig, tables, sizeof(double) * (size_t)m->toff[m->nc]);
  memcpy(m->prior, prior, sizeof(double) * (size_t)m->prior_off[m->nv]);
  orc_reset_model(m);
}

/* ---- slice-to-slice messages ------------------------------------------- */
static int iface_size(const orc_model* m) {
  int i, n = 1;
  for (i = 0; i < m->nif; i++) n *= m->card[m->outg[i]];
  return n;
}

/* start_timeslice_message_pass, src/nip.c:1031-1065: normalised marginal of
 * out_clique over I_t (forward) or of in_clique over I_{t-1} (backward) */
static void start_message(orc_model* m, int forward, double* msg) {
  int map[MAXD], ccard[MAXD], icard[MAXD], c;
  const int* vars;
  if (m->nif == 0) { msg[0] = 1.0; return; }
  vars = forward ? m->outg : m->prev;
  c = forward ? m->out_clique : m->in_clique;
  cards_of(m, cvars(m, c), cdim(m, c), ccard);
  cards_of(m, m->outg, m->nif, icard);
  mapper(cvars(m, c), cdim(m, c), vars, m->nif, map);
  orc_general_marginalise(m->p + m->toff[c], cdim(m, c), ccard, msg, m->nif, icard, map);
  orc_normalise_array(msg, iface_size(m));
}

/* finish_timeslice_message_pass, src/nip.c:1069-1098 */
static void finish_message(orc_model* m, int forward, const double* num, const double* den) {
  int map[MAXD], ccard[MAXD], icard[MAXD], c;
  const int* vars;
  if (m->nif == 0) return;
  vars = forward ? m->prev : m->outg;
  c = forward ? m->in_clique : m->out_clique;
  cards_of(m, cvars(m, c), cdim(m, c), ccard);
  cards_of(m, m->outg, m->nif, icard);
  mapper(cvars(m, c), cdim(m, c), vars, m->nif, map);
  orc_update_potential(num, den, m->nif, icard, m->p + m->toff[c], cdim(m, c), ccard, map);
}

/* insert_ts_step, src/nip.c:982-1001; `mask[v]` carries the NIP_MARK bits */
static void insert_step(orc_model* m, int n_obs, const int* obs_vars, const int* row,
                        const unsigned char* selected) {
  int k;
  for (k = 0; k < n_obs; k++)
    if ((!selected || selected[obs_vars[k]]) && row[k] >= 0)
      orc_enter_index_observation(m, obs_vars[k], row[k]);
}

static void write_marginals(orc_model* m, int nq, const int* q, double* out) {
  int i;
  for (i = 0; i < nq; i++) { orc_marginal(m, q[i], out); out += m->card[q[i]]; }
}

/* forward_inference (src/nip.c:1103-1315) / forward_backward_inference
 * (src/nip.c:1320-1581) */
int orc_infer(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
              const unsigned char* use_evidence, int nq, const int* q, int forward_only,
              int want_ll, double* post, double* loglik) {
  int isz = iface_size(m), t, i, row = 0;
  double** ag = (double**)calloc((size_t)T + 1, sizeof(double*));
  double m1 = 0, m2, ll = 0;
  for (i = 0; i < nq; i++) row += m->card[q[i]];
  for (t = 0; t <= T; t++) {
    ag[t] = (double*)calloc((size_t)isz, sizeof(double));
    for (i = 0; i < isz; i++) ag[t][i] = 1.0;
  }
  orc_reset_model(m);
  orc_use_priors(m, 0);
  for (t = 0; t < T; t++) { /* forward phase, :1435-1493 */
    if (t > 0) finish_message(m, 1, ag[t - 1], NULL);
    if (want_ll) { orc_make_consistent(m); m1 = orc_prob_mass(m); }
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, use_evidence);
    orc_make_consistent(m);
    if (want_ll) {
      m2 = orc_prob_mass(m);
      if (m1 > 0 && m2 > 0) ll += log(m2) - log(m1);
      if (m2 == 0) ll = -DBL_MAX;
    }
    if (forward_only && post) write_marginals(m, nq, q, post + (size_t)t * row); /* :1273-1289 */
    start_message(m, 1, ag[t]);
    orc_reset_model(m);
    if (forward_only) orc_use_priors(m, 1);      /* :1309-1310 */
    else orc_use_priors(m, T > 1 ? 1 : 0);       /* :1488-1492 */
  }
  for (t = T - 1; !forward_only && t >= 0; t--) { /* backward phase, :1498-1573 */
    if (t > 0) finish_message(m, 1, ag[t - 1], NULL);
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, use_evidence);
    if (t < T - 1) finish_message(m, 0, ag[t + 1], ag[t]);
    orc_make_consistent(m);
    if (post) write_marginals(m, nq, q, post + (size_t)t * row);
    if (t > 0) start_message(m, 0, ag[t]);
    orc_reset_model(m);
    orc_use_priors(m, t > 1 ? 1 : 0);
  }
  for (t = 0; t <= T; t++) free(ag[t]);
  free(ag);
  if (loglik && want_ll) *loglik = ll;
  return 0;
}

/* nip_find_family_mapping, src/nipjointree.c:1001-1040: child first, then
 * parents[] order */
static int family_mapping(const orc_model* m, int v, int* map, int* fcard) {
  int c = m->family[v], j, n = 1 + m->poff[v + 1] - m->poff[v];
  map[0] = var_pos_in_clique(m, c, v);
  fcard[0] = m->card[v];
  for (j = 1; j < n; j++) {
    int par = m->parents[m->poff[v] + j - 1];
    map[j] = var_pos_in_clique(m, c, par);
    fcard[j] = m->card[par];
  }
  return n;
}

/* e_step, src/nip.c:1708-2007 */
int orc_estep(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
              const unsigned char* use_evidence, double* counts, double* loglik) {
  int isz = iface_size(m), t, i, v, status = 0;
  double** ag = (double**)calloc((size_t)T + 1, sizeof(double*));
  double m1, m2, ll = 0;
  double* fam = (double*)calloc((size_t)orc_counts_size(m), sizeof(double));
  for (t = 0; t <= T; t++) {
    ag[t] = (double*)calloc((size_t)isz, sizeof(double));
    for (i = 0; i < isz; i++) ag[t][i] = 1.0;
  }
  orc_reset_model(m);
  orc_use_priors(m, 0);
  for (t = 0; t < T; t++) { /* :1791-1880 */
    if (t > 0) finish_message(m, 1, ag[t - 1], NULL);
    orc_make_consistent(m);
    m1 = orc_prob_mass(m);
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, use_evidence);
    orc_make_consistent(m);
    m2 = orc_prob_mass(m);
    if (m1 > 0 && m2 > 0) ll += log(m2) - log(m1);
    if (m1 <= 0 || m2 <= 0 || ll > 0) { status = NIPGPU_EBADLUCK; goto out; } /* :1827-1854 */
    start_message(m, 1, ag[t]);
    orc_reset_model(m);
    orc_use_priors(m, T > 1 ? 1 : 0);
  }
  for (t = T - 1; t >= 0; t--) { /* :1885-1990 */
    if (t > 0) finish_message(m, 1, ag[t - 1], NULL);
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, use_evidence);
    if (t < T - 1) finish_message(m, 0, ag[t + 1], ag[t]);
    orc_make_consistent(m);
    for (v = 0; v < m->nv; v++) {
      int map[MAXD], fcard[MAXD], ccard[MAXD], n, c = m->family[v];
      long sz = m->coff[v + 1] - m->coff[v], k;
      if (t > 0 && (m->flags[v] & NIPGPU_IF_OLD_OUTGOING)) continue; /* :1932 */
      n = family_mapping(m, v, map, fcard);
      cards_of(m, cvars(m, c), cdim(m, c), ccard);
      orc_general_marginalise(m->p + m->toff[c], cdim(m, c), ccard, fam + m->coff[v], n, fcard, map);
      orc_normalise_array(fam + m->coff[v], (int)sz); /* :1962 whole table sums to 1 */
      for (k = 0; k < sz; k++) counts[m->coff[v] + k] += fam[m->coff[v] + k]; /* :1965 */
    }
    if (t > 0) start_message(m, 0, ag[t]);
    orc_reset_model(m);
    orc_use_priors(m, t > 1 ? 1 : 0);
  }
out:
  for (t = 0; t <= T; t++) free(ag[t]);
  free(ag);
  free(fam);
  if (loglik) *loglik = ll;
  return status;
}

/* m_step, src/nip.c:2010-2071 */
int orc_mstep(orc_model* m, double* counts) {
  int v, j;
  for (v = 0; v < m->nv; v++)
    orc_normalise_cpd(counts + m->coff[v], (int)(m->coff[v + 1] - m->coff[v]), m->card[v]);
  orc_total_reset(m);
  for (v = 0; v < m->nv; v++) {
    if (m->poff[v + 1] > m->poff[v]) {
      int map[MAXD], fcard[MAXD], ccard[MAXD], c = m->family[v];
      int n = family_mapping(m, v, map, fcard);
      cards_of(m, cvars(m, c), cdim(m, c), ccard);
      orc_init_potential(counts + m->coff[v], n, fcard, m->p + m->toff[c], cdim(m, c), ccard, map);
      orc_init_potential(counts + m->coff[v], n, fcard, m->orig + m->toff[c], cdim(m, c), ccard, map);
    } else
      for (j = 0; j < m->card[v]; j++) m->prior[m->prior_off[v] + j] = counts[m->coff[v] + j];
  }
  return 0;
}

/* util/niplikelihood.c:111-135 — slices are evaluated independently, only the
 * prior rule distinguishes t == 0 from t > 0 */
int orc_likelihood(orc_model* m, int T, int n_obs, const int* obs_vars, const int* data,
                   const unsigned char* evidence_off, const unsigned char* evidence_on,
                   double* out) {
  int t;
  orc_reset_model(m);
  orc_use_priors(m, 0);
  for (t = 0; t < T; t++) {
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, evidence_off);
    orc_make_consistent(m);
    out[2 * t] = orc_prob_mass(m);
    insert_step(m, n_obs, obs_vars, data + (size_t)t * n_obs, evidence_on);
    orc_make_consistent(m);
    out[2 * t + 1] = orc_prob_mass(m);
    orc_reset_model(m);
    orc_use_priors(m, 1);
  }
  return 0;
}
