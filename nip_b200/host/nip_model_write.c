/* nip_model_write.c — a lossless companion of write_model() (src/nip.c:298-482).
 *
 * The reference's writer prints probabilities with "%f" (six decimals, src/nip.c:388,462),
 * attaches NIP_next to the wrong end of the slice link, and keeps the declaration order, which
 * makes its own parser normalise along a parent axis on reload (it normalises storage dimension 0,
 * src/huginnet.y:635-636).  A model trained on the GPU backend therefore changes when it goes
 * through a file.  nip_gpu_write_model_exact() writes the same Hugin .net text with
 *   - "%.17g" numbers (doubles round-trip),
 *   - every node declared before its parents, so that the child is dimension 0 of each potential
 *     the parser builds,
 *   - NIP_next on the previous-slice variable, pointing to its successor,
 * so that parse_model() of the written file gives the same conditional distributions
 * (tests/test_host_io.py).  Host-side only; no device code involved. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nip.h"
#include "nipjointree.h"
#include "nippotential.h"
#include "nipvariable.h"

#ifndef NIP_ERROR_GENERAL
#define NIP_ERROR_NULLPOINTER 1
#define NIP_ERROR_OUTOFMEMORY 4
#define NIP_ERROR_IO 5
#endif

#define VALUES_PER_LINE 6

static void put_values(FILE* f, const double* v, int n, int block) {
  int j;
  for (j = 0; j < n; j++) {
    if (j > 0 && (j % block == 0 || (j % block) % VALUES_PER_LINE == 0)) fputs("\n            ", f);
    fprintf(f, " %.17g", v[j]);
  }
}

int nip_gpu_write_model_exact(nip_model model, char* filename) {
  FILE* f;
  int i, j, k, n, x, y, done = 0;
  int* emitted;
  if (!model || !filename) return NIP_ERROR_NULLPOINTER;
  n = model->num_of_vars;
  emitted = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  if (!emitted) return NIP_ERROR_OUTOFMEMORY;
  f = fopen(filename, "w");
  if (!f) { free(emitted); return NIP_ERROR_IO; }
  fprintf(f, "net\n{\n    node_size = (%d %d);\n}\n", model->node_size_x, model->node_size_y);

  /* nodes: a variable is declared once every variable it is a parent of has been declared */
  while (done < n) {
    int progressed = 0;
    for (i = 0; i < n; i++) {
      nip_variable v = model->variables[i];
      int ready = !emitted[i];
      for (j = 0; j < n && ready; j++) {
        nip_variable c = model->variables[j];
        if (emitted[j]) continue;
        for (k = 0; k < nip_number_of_parents(c); k++)
          if (c->parents[k] == v && c != v) ready = 0;
      }
      if (!ready) continue;
      nip_get_variable_position(v, &x, &y);
      fprintf(f, "\nnode %s\n{\n    label = \"%s\";\n    position = (%d %d);\n    states = (", nip_variable_symbol(v),
              v->name ? v->name : "", x, y);
      for (j = 0; j < NIP_CARDINALITY(v); j++) fprintf(f, " \"%s\"", v->state_names[j]);
      fputs(" );\n", f);
      if (v->next) fprintf(f, "    NIP_next = \"%s\";\n", nip_variable_symbol(v->next));
      fputs("}\n", f);
      emitted[i] = 1;
      done++;
      progressed = 1;
    }
    if (!progressed) break;  /* a cycle: cannot happen in a parsed model */
  }
  free(emitted);
  if (done < n) { fclose(f); return NIP_ERROR_IO; }

  /* priors of the independent variables */
  for (i = 0; i < model->num_of_vars - model->num_of_children; i++) {
    nip_variable v = model->independent[i];
    fprintf(f, "\npotential (%s)\n{\n    data = (", nip_variable_symbol(v));
    if (v->prior) put_values(f, v->prior, NIP_CARDINALITY(v), NIP_CARDINALITY(v));
    else for (j = 0; j < NIP_CARDINALITY(v); j++) fputs(" 1", f);
    fputs(" );\n}\n", f);
  }

  /* conditional distributions: the family marginal of original_p, normalised over the child,
   * exactly as write_model() forms them (src/nip.c:420-437) */
  for (i = 0; i < model->num_of_children; i++) {
    nip_variable v = model->children[i];
    const int np = nip_number_of_parents(v);
    int* card = (int*)calloc((size_t)np + 1, sizeof(int));
    nip_potential p;
    nip_clique c;
    if (!card) { fclose(f); return NIP_ERROR_OUTOFMEMORY; }
    fprintf(f, "\npotential (%s |", nip_variable_symbol(v));
    for (j = np - 1; j >= 0; j--) fprintf(f, " %s", nip_variable_symbol(v->parents[j]));  /* Hugin: reverse order */
    fputs(")\n{\n    data = (", f);
    card[0] = NIP_CARDINALITY(v);
    for (j = 0; j < np; j++) card[j + 1] = NIP_CARDINALITY(v->parents[j]);
    p = nip_new_potential(card, np + 1, NULL);
    c = nip_find_family(model->cliques, model->num_of_cliques, v);
    if (!p || !c) { free(card); fclose(f); return NIP_ERROR_OUTOFMEMORY; }
    nip_general_marginalise(c->original_p, p, nip_find_family_mapping(c, v));
    nip_normalise_cpd(p);
    put_values(f, p->data, p->size_of_data, NIP_CARDINALITY(v));
    fputs(" );\n}\n", f);
    nip_free_potential(p);
    free(card);
  }
  return fclose(f) ? NIP_ERROR_IO : NIP_NO_ERROR;
}
