/* TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
 *
 * Hand-written recursive-descent reader for the Hugin .net subset that NIP
 * accepts.  The reference generates its reader from src/huginnet.y with GNU
 * Bison, which is not available in this image, so the `oracle/_ref` build
 * links this file in its place.  It exports the six entry points that
 * src/nip.c:39-43 expects from the generated parser
 *   yyparse, open_net_file, close_net_file,
 *   get_parsed_variables, get_cliques, get_parsed_node_size
 * and performs the same sequence of library calls as the grammar actions
 * (src/huginnet.y:202-235 input, :318-383 node, :570-650 potential,
 *  :1049-1098 graph, :1110-1152 potentials→join tree, :1155-1254 interface),
 * so that the join tree handed to the hot path is the one HEAD would build:
 *   - variables are created in declaration order (ids ascend),
 *   - the parents named after '|' are PREPENDED (:758) i.e. reversed,
 *   - every parsed CPT is normalised along storage dimension 0 (:635-636),
 *   - priors of parentless nodes go to variable->prior, not into cliques.
 * Only model construction lives here; all inference arithmetic stays in the
 * reference's own nippotential.c / nipjointree.c / nip.c.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "niplists.h"
#include "nipgraph.h"
#include "nipparsers.h"
#include "nipjointree.h"
#include "nipvariable.h"
#include "nippotential.h"
#include "niperrorhandler.h"

/* ---- token classes ---------------------------------------------------- */
enum tk {
  TK_EOF = 0, TK_CHAR, TK_QSTR, TK_USTR, TK_NUM,
  TK_NET, TK_CLASS, TK_NODE_SIZE, TK_DATA, TK_UTILITY, TK_DECISION,
  TK_DISCRETE, TK_CONTINUOUS, TK_NODE, TK_LABEL, TK_POSITION, TK_STATES,
  TK_NEXT, TK_POTENTIAL, TK_NORMAL
};

static const struct { const char* word; enum tk kind; } keywords[] = {
  {"net", TK_NET}, {"class", TK_CLASS}, {"node_size", TK_NODE_SIZE},
  {"data", TK_DATA}, {"utility", TK_UTILITY}, {"decision", TK_DECISION},
  {"discrete", TK_DISCRETE}, {"continuous", TK_CONTINUOUS}, {"node", TK_NODE},
  {"label", TK_LABEL}, {"position", TK_POSITION}, {"states", TK_STATES},
  {"NIP_next", TK_NEXT}, {"potential", TK_POTENTIAL}, {"normal", TK_NORMAL},
};

/* ---- reader state (the generated parser is just as global) ------------ */
static FILE* rd_file = NULL;
static int rd_file_is_open = 0;

static enum tk cur_kind;     /* look-ahead token */
static int cur_char;         /* for TK_CHAR */
static char* cur_text;       /* for TK_QSTR / TK_USTR (owned) */
static double cur_num;       /* for TK_NUM */

static nip_variable_list all_vars = NULL;
static nip_potential_list cpt_list = NULL;
static nip_interface_list next_links = NULL;
static nip_clique* jt_cliques = NULL;
static int jt_ncliques = 0;
static int net_node_w = 80, net_node_h = 60;

static void advance(void) {
  int len = 0;
  char* raw;
  size_t k;
  free(cur_text);
  cur_text = NULL;
  raw = nip_next_hugin_token(rd_file, &len);
  if (len <= 0 || !raw) { cur_kind = TK_EOF; free(raw); return; }
  if (len == 1) {
    unsigned char c = (unsigned char)raw[0];
    if (isalpha(c)) { cur_kind = TK_USTR; cur_text = raw; return; }
    if (isdigit(c)) { cur_kind = TK_NUM; cur_num = (double)(c - '0'); free(raw); return; }
    cur_kind = TK_CHAR; cur_char = c; free(raw); return;
  }
  for (k = 0; k < sizeof(keywords) / sizeof(keywords[0]); k++)
    if ((int)strlen(keywords[k].word) == len && strcmp(keywords[k].word, raw) == 0) {
      cur_kind = keywords[k].kind; free(raw); return;
    }
  if (raw[0] == '"' && raw[len - 1] == '"') {
    memmove(raw, raw + 1, (size_t)len - 2);
    raw[len - 2] = '\0';
    cur_kind = TK_QSTR; cur_text = raw; return;
  }
  {
    char* end = NULL;
    double x = strtod(raw, &end);
    if (!(end == raw && x == 0)) { cur_kind = TK_NUM; cur_num = x; free(raw); return; }
  }
  cur_kind = TK_USTR; cur_text = raw;
}

static int is_char(int c) { return cur_kind == TK_CHAR && cur_char == c; }

static int eat_char(int c) {
  if (!is_char(c)) {
    fprintf(stderr, "NET reader: expected '%c'\n", c);
    return 0;
  }
  advance();
  return 1;
}

static char* take_text(void) { char* t = cur_text; cur_text = NULL; return t; }

/* ( NUMBER NUMBER ) ; */
static int read_pair(int* a, int* b) {
  if (!eat_char('(')) return 0;
  if (cur_kind != TK_NUM) return 0;
  *a = abs((int)cur_num); advance();
  if (cur_kind != TK_NUM) return 0;
  *b = abs((int)cur_num); advance();
  return eat_char(')') && eat_char(';');
}

/* numbers := ( NUMBER | '(' numbers ')' )*   — flattens nesting */
static int read_numbers(nip_double_list out) {
  for (;;) {
    if (cur_kind == TK_NUM) {
      if (out && nip_append_double(out, cur_num) != NIP_NO_ERROR) return 0;
      advance();
    } else if (is_char('(')) {
      advance();
      if (!read_numbers(out)) return 0;
      if (!eat_char(')')) return 0;
    } else
      return 1;
  }
}

/* UNQUOTED '=' value ';'   (ignored field) — cur is the UNQUOTED name */
static int skip_unknown_field(void) {
  advance();
  if (!eat_char('=')) return 0;
  if (cur_kind == TK_QSTR) advance();
  else if (!read_numbers(NULL)) return 0;
  return eat_char(';');
}

static int read_net_parameters(void) {
  for (;;) {
    if (cur_kind == TK_NODE_SIZE) {
      advance();
      if (!eat_char('=')) return 0;
      if (!read_pair(&net_node_w, &net_node_h)) return 0;
    } else if (cur_kind == TK_USTR) {
      if (!skip_unknown_field()) return 0;
    } else
      return 1;
  }
}

static void free_strings(char** s, int n) {
  int i;
  if (!s) return;
  for (i = 0; i < n; i++) free(s[i]);
  free(s);
}

/* node NAME { fields }  (optionally prefixed by 'discrete') */
static int read_node(void) {
  char *symbol = NULL, *label = NULL, *next_name = NULL;
  char** states = NULL;
  int nstates = 0, px = 100, py = 100, ok = 0;
  nip_variable v;

  if (cur_kind == TK_CONTINUOUS || cur_kind == TK_UTILITY || cur_kind == TK_DECISION) {
    fprintf(stderr, "NET reader: continuous / utility / decision nodes are not supported.\n");
    nip_report_error(__FILE__, __LINE__, ENOSYS, 1);
    return 0;
  }
  if (cur_kind == TK_DISCRETE) advance();
  if (cur_kind != TK_NODE) return 0;
  advance();
  if (cur_kind != TK_USTR) return 0;
  symbol = take_text();
  advance();
  if (!eat_char('{')) goto done;

  for (;;) {
    if (cur_kind == TK_STATES) {
      nip_string_list names = nip_new_string_list();
      advance();
      if (!eat_char('=') || !eat_char('(')) { free(names); goto done; }
      while (cur_kind == TK_QSTR) {
        nip_append_string(names, take_text());
        advance();
      }
      free_strings(states, nstates);
      states = nip_string_list_to_array(names);
      nstates = NIP_LIST_LENGTH(names);
      nip_empty_string_list(names);
      free(names);
      if (!eat_char(')') || !eat_char(';')) goto done;
    } else if (cur_kind == TK_LABEL || cur_kind == TK_NEXT) {
      int is_label = (cur_kind == TK_LABEL);
      advance();
      if (!eat_char('=') || cur_kind != TK_QSTR) goto done;
      if (is_label) { free(label); label = take_text(); }
      else { free(next_name); next_name = take_text(); }
      advance();
      if (!eat_char(';')) goto done;
    } else if (cur_kind == TK_POSITION) {
      advance();
      if (!eat_char('=') || !read_pair(&px, &py)) goto done;
    } else if (cur_kind == TK_USTR) {
      if (!skip_unknown_field()) goto done;
    } else
      break;
  }
  if (!eat_char('}')) goto done;

  if (!states) {
    fprintf(stderr, "NET reader: The states field is missing (node %s)\n", symbol);
    nip_report_error(__FILE__, __LINE__, EINVAL, 1);
    goto done;
  }
  v = nip_new_variable(symbol, label ? label : " ", states, nstates);
  if (!v) { nip_report_error(__FILE__, __LINE__, EINVAL, 1); goto done; }
  nip_set_variable_position(v, px, py);
  if (!all_vars) all_vars = nip_new_variable_list();
  nip_append_variable(all_vars, v);
  if (next_name) {
    if (!next_links) next_links = nip_new_interface_list();
    if (nip_append_interface(next_links, v, next_name) != NIP_NO_ERROR) goto done;
    next_name = NULL; /* the list owns the string now */
  }
  ok = 1;
done:
  free(symbol); free(label); free(next_name);
  free_strings(states, nstates);
  return ok;
}

/* potential ( CHILD [ '|' PARENT* ] ) { [ data = ( numbers ) ; ] } */
static int read_potential(void) {
  nip_variable child;
  nip_variable_list rev_parents = NULL; /* built by prepending, as HEAD does */
  nip_variable* parents = NULL;
  nip_variable* family = NULL;
  double* values = NULL;
  int nvalues = 0, nparents = 0, conditional = 0, i, size, ok = 0;
  nip_potential p;

  advance(); /* 'potential' */
  if (!eat_char('(') || cur_kind != TK_USTR) return 0;
  child = nip_search_variable_list(all_vars, cur_text);
  advance();
  if (is_char('|')) {
    conditional = 1;
    advance();
    rev_parents = nip_new_variable_list();
    while (cur_kind == TK_USTR) {
      if (nip_prepend_variable(rev_parents, nip_search_variable_list(all_vars, cur_text))
          != NIP_NO_ERROR) {
        nip_report_error(__FILE__, __LINE__, EINVAL, 1);
        goto done;
      }
      advance();
    }
  }
  if (!eat_char(')') || !eat_char('{')) goto done;
  if (cur_kind == TK_DATA) {
    nip_double_list nums = nip_new_double_list();
    advance();
    if (!eat_char('=') || !eat_char('(') || !read_numbers(nums) ||
        !eat_char(')') || !eat_char(';')) {
      nip_empty_double_list(nums); free(nums); goto done;
    }
    values = nip_double_list_to_array(nums);
    nvalues = NIP_LIST_LENGTH(nums);
    nip_empty_double_list(nums);
    free(nums);
    if (!values) { nip_report_error(__FILE__, __LINE__, EINVAL, 1); goto done; }
  }
  if (!eat_char('}')) goto done;
  if (!child) { nip_report_error(__FILE__, __LINE__, EINVAL, 1); goto done; }

  if (conditional) {
    nparents = NIP_LIST_LENGTH(rev_parents);
    if (nparents == 0) goto done; /* "potential (A | )" aborts in HEAD too */
    parents = nip_variable_list_to_array(rev_parents);
  }
  family = (nip_variable*)calloc((size_t)nparents + 1, sizeof(nip_variable));
  if (!family) goto done;
  family[0] = child;
  size = NIP_CARDINALITY(child);
  for (i = 0; i < nparents; i++) {
    family[i + 1] = parents[i];
    size *= NIP_CARDINALITY(parents[i]);
  }
  if (values && size > nvalues) {
    fprintf(stderr, "NET reader: Not enough elements in potential( %s... )!\n",
            nip_variable_symbol(child));
    goto done;
  }
  if (!cpt_list) cpt_list = nip_new_potential_list();
  p = nip_create_potential(family, nparents + 1, values);
  if (conditional) nip_normalise_cpd(p);           /* dimension 0 = lowest id */
  else if (values) nip_normalise_potential(p);
  if (nip_append_potential(cpt_list, p, child, parents) != NIP_NO_ERROR) {
    nip_report_error(__FILE__, __LINE__, EINVAL, 1);
    goto done;
  }
  parents = NULL; /* owned by the list */
  ok = 1;
done:
  if (rev_parents) { nip_empty_variable_list(rev_parents); free(rev_parents); }
  free(parents); free(family); free(values);
  return ok;
}

/* NIP_next links → variable->next / previous and the interface flags */
static int link_time_slices(void) {
  nip_interface_link l;
  nip_variable_iterator it;
  nip_variable v, u;
  int i, has_old_parent;

  if (!next_links) return NIP_NO_ERROR;
  for (l = NIP_LIST_ITERATOR(next_links); l; l = NIP_LIST_NEXT(l)) {
    v = l->var;
    u = nip_search_variable_list(all_vars, l->next);
    if (!u || NIP_CARDINALITY(u) != NIP_CARDINALITY(v)) {
      fprintf(stderr, "NET reader: Invalid 'NIP_next' field for node %s.\n",
              nip_variable_symbol(v));
      return nip_report_error(__FILE__, __LINE__, EINVAL, 1);
    }
    v->next = u;      /* u lives in slice t   */
    u->previous = v;  /* v lives in slice t-1 */
  }
  it = NIP_LIST_ITERATOR(all_vars);
  while ((u = nip_next_variable(&it)) != NULL) {
    has_old_parent = 0;
    for (i = 0; i < nip_number_of_parents(u); i++) {
      v = u->parents[i];
      if (v->next) {
        v->interface_status |= NIP_INTERFACE_OLD_OUTGOING;
        v->next->interface_status |= NIP_INTERFACE_OUTGOING;
        u->interface_status |= NIP_INTERFACE_INCOMING;
        has_old_parent = 1;
      }
    }
    if (has_old_parent)
      for (i = 0; i < nip_number_of_parents(u); i++)
        if (u->parents[i]->next == NULL)
          u->parents[i]->interface_status |= NIP_INTERFACE_INCOMING;
  }
  return NIP_NO_ERROR;
}

static int build_join_tree(void) {
  nip_graph g;
  nip_variable_iterator it;
  nip_variable v;
  nip_potential_link l;
  int i, nparents;

  if (!all_vars) return EINVAL;
  g = nip_new_graph((unsigned)NIP_LIST_LENGTH(all_vars));
  it = NIP_LIST_ITERATOR(all_vars);
  while ((v = nip_next_variable(&it)) != NULL)
    if (nip_graph_add_node(g, v) != NIP_NO_ERROR) return EINVAL;

  for (l = cpt_list ? NIP_LIST_ITERATOR(cpt_list) : NULL; l; l = NIP_LIST_NEXT(l)) {
    nparents = NIP_DIMENSIONALITY(l->data) - 1;
    for (i = 0; i < nparents; i++)
      if (nip_graph_add_child(g, l->parents[i], l->child) != NIP_NO_ERROR) return EINVAL;
    nip_set_parents(l->child, l->parents, nparents);
  }

  if (link_time_slices() != NIP_NO_ERROR) {
    fprintf(stderr, "Invalid timeslice specification!\nCheck NIP_next declarations.\n");
    return EINVAL;
  }
  if (next_links) { nip_free_interface_list(next_links); next_links = NULL; }

  jt_ncliques = nip_graph_to_cliques(g, &jt_cliques);
  nip_free_graph(g);
  if (jt_ncliques < 0) return EINVAL;

  for (l = cpt_list ? NIP_LIST_ITERATOR(cpt_list) : NULL; l; l = NIP_LIST_NEXT(l)) {
    nip_clique home = nip_find_family(jt_cliques, jt_ncliques, l->child);
    if (!home) { fprintf(stderr, "NET reader: no family clique for %s\n", l->child->symbol); continue; }
    if (NIP_DIMENSIONALITY(l->data) > 1) {
      if (nip_init_clique(home, l->child, l->data, 0) != NIP_NO_ERROR) return EINVAL;
    } else if (nip_set_prior(l->child, l->data->data) != NIP_NO_ERROR)
      return EINVAL;
  }
  if (cpt_list) { nip_free_potential_list(cpt_list); cpt_list = NULL; }
  return NIP_NO_ERROR;
}

/* ---- the six entry points src/nip.c links against ---------------------- */
int yyparse(void) {
  int wrapped_in_class = 0;
  cur_text = NULL;
  jt_cliques = NULL; jt_ncliques = 0;
  advance();
  if (cur_kind == TK_NET) {
    advance();
    if (!eat_char('{') || !read_net_parameters() || !eat_char('}')) return 1;
  } else if (cur_kind == TK_CLASS) {
    advance();
    if (cur_kind != TK_USTR) return 1;
    advance();
    if (!eat_char('{') || !read_net_parameters()) return 1;
    wrapped_in_class = 1;
  }
  while (cur_kind == TK_NODE || cur_kind == TK_DISCRETE || cur_kind == TK_CONTINUOUS ||
         cur_kind == TK_UTILITY || cur_kind == TK_DECISION)
    if (!read_node()) return 1;
  while (cur_kind == TK_POTENTIAL)
    if (!read_potential()) return 1;
  if (wrapped_in_class && !eat_char('}')) return 1;
  if (cur_kind != TK_EOF) {
    fprintf(stderr, "NET reader: syntax error\n");
    /* drain so that the tokeniser's static line state is clean for the next file */
    while (cur_kind != TK_EOF) advance();
    return 1;
  }
  if (build_join_tree() != NIP_NO_ERROR) {
    nip_report_error(__FILE__, __LINE__, EINVAL, 1);
    return 1;
  }
  return 0;
}

FILE* open_net_file(const char* filename) {
  if (!rd_file_is_open) {
    rd_file = fopen(filename, "r");
    if (!rd_file) { nip_report_error(__FILE__, __LINE__, EIO, 1); return NULL; }
    rd_file_is_open = 1;
  }
  return rd_file;
}

void close_net_file(void) {
  if (rd_file_is_open) { fclose(rd_file); rd_file_is_open = 0; rd_file = NULL; }
}

nip_variable_list get_parsed_variables(void) { return all_vars; }

int get_cliques(nip_clique** out) { *out = jt_cliques; return jt_ncliques; }

void get_parsed_node_size(int* x, int* y) { *x = net_node_w; *y = net_node_h; }
