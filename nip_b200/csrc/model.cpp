// model.cpp — see model.h
#include "model.h"

#include <algorithm>
#include <functional>
#include <set>

namespace nipgpu {

int HostModel::var_pos(int c, int v) const {
  for (int k = 0; k < clique_dim(c); k++)
    if (clique_vars(c)[k] == v) return k;
  return -1;
}

static int pow2_floor(int x) {
  int p = 1;
  while (p * 2 <= x) p *= 2;
  return p;
}

int HostModel::add_proj(int clique, const std::vector<int>& vars) {
  for (size_t i = 0; i < projs.size(); i++)
    if (projs[i].clique == clique && projs[i].vars == vars) return (int)i;
  Proj p;
  p.clique = clique;
  p.vars = vars;
  const int nd = clique_dim(clique);
  const int* cv = clique_vars(clique);
  std::vector<int64_t> stride(nd);
  int64_t st = 1;
  for (int k = 0; k < nd; k++) { stride[k] = st; st *= card[cv[k]]; }
  std::vector<int> dpos;  // position of every destination variable in the clique
  for (int v : vars) dpos.push_back(var_pos(clique, v));
  p.m = 1;
  for (int v : vars) p.m *= card[v];
  p.base.resize(p.m);
  for (int j = 0; j < p.m; j++) {
    int rem = j;
    int64_t b = 0;
    for (size_t k = 0; k < vars.size(); k++) {
      b += (int64_t)(rem % card[vars[k]]) * stride[dpos[k]];
      rem /= card[vars[k]];
    }
    p.base[j] = (int)b;
  }
  std::vector<int> freepos;
  for (int k = 0; k < nd; k++)
    if (std::find(dpos.begin(), dpos.end(), k) == dpos.end()) freepos.push_back(k);
  p.R = 1;
  for (int k : freepos) p.R *= card[cv[k]];
  p.off.resize(p.R);
  for (int r = 0; r < p.R; r++) {
    int rem = r;
    int64_t o = 0;
    for (int k : freepos) {
      o += (int64_t)(rem % card[cv[k]]) * stride[k];
      rem /= card[cv[k]];
    }
    p.off[r] = (int)o;
  }
  // Thread mapping: if the destination holds the clique's fastest dimension,
  // consecutive destination entries are consecutive addresses -> one thread per
  // destination entry; otherwise the free offsets are the contiguous ones ->
  // up to a warp cooperates on one destination entry (shuffle reduction).
  const bool dest_has_dim0 = std::find(dpos.begin(), dpos.end(), 0) != dpos.end();
  p.lanes = dest_has_dim0 ? 1 : std::min(32, pow2_floor(std::max(1, p.R)));
  // table-order map, split after the leading dimensions whose product reaches 1024
  std::vector<int64_t> dstride(nd, 0);  // weight of clique dimension k inside j (0: summed out)
  {
    int64_t w = 1;
    for (size_t k = 0; k < vars.size(); k++) { dstride[dpos[k]] = w; w *= card[vars[k]]; }
  }
  int split = 0;
  p.F = 1;
  while (split < nd && p.F < 1024) p.F *= card[cv[split++]];
  auto fill = [&](std::vector<int>& out, int from, int to) {
    int64_t count = 1;
    for (int k = from; k < to; k++) count *= card[cv[k]];
    out.resize((size_t)count);
    for (int64_t x = 0; x < count; x++) {
      int64_t rem = x, j = 0;
      for (int k = from; k < to; k++) { j += (rem % card[cv[k]]) * dstride[k]; rem /= card[cv[k]]; }
      out[(size_t)x] = (int)j;
    }
  };
  fill(p.jlo, 0, split);
  fill(p.jhi, split, nd);
  projs.push_back(p);
  return (int)projs.size() - 1;
}

template <class T, class U>
static std::vector<T> copy_n_(const U* a, int64_t n) {
  std::vector<T> r((size_t)std::max<int64_t>(n, 0));
  for (int64_t i = 0; i < n; i++) r[(size_t)i] = (T)a[i];
  return r;
}

std::string HostModel::load(const nipgpu_model_desc* d) {
  if (!d) return "null description";
  nv = d->n_vars; nc = d->n_cliques; ns = d->n_sepsets; nif = d->n_interface;
  in_clique = d->in_clique; out_clique = d->out_clique;
  if (nv <= 0 || nc <= 0) return "model needs at least one variable and one clique";
  card = copy_n_<int>(d->var_card, nv);
  flags = copy_n_<int>(d->var_flags, nv);
  poff = copy_n_<int>(d->var_parent_off, nv + 1);
  parents = copy_n_<int>(d->var_parents, poff[nv]);
  family = copy_n_<int>(d->var_family, nv);
  prior_off = copy_n_<int>(d->var_prior_off, nv + 1);
  prior = copy_n_<double>(d->var_prior, prior_off[nv]);
  cvoff = copy_n_<int>(d->clique_var_off, nc + 1);
  cvars = copy_n_<int>(d->clique_vars, cvoff[nc]);
  toff = copy_n_<int64_t>(d->clique_tab_off, nc + 1);
  tables = copy_n_<double>(d->clique_tables, toff[nc]);
  scl = copy_n_<int>(d->sepset_cliques, 2 * (int64_t)ns);
  svoff = copy_n_<int>(d->sepset_var_off, ns + 1);
  svars = copy_n_<int>(d->sepset_vars, svoff[ns]);
  adjoff = copy_n_<int>(d->clique_adj_off, nc + 1);
  adj = copy_n_<int>(d->clique_adj, adjoff[nc]);
  outg = copy_n_<int>(d->outgoing, nif);
  prev = copy_n_<int>(d->prev_outgoing, nif);

  // ---- validation ----
  for (int v = 0; v < nv; v++) {
    if (card[v] <= 0) return "variable with non-positive cardinality";
    if (family[v] < 0 || family[v] >= nc) return "family clique out of range";
    if (var_pos(family[v], v) < 0) return "family clique does not hold its variable";
    for (int j = poff[v]; j < poff[v + 1]; j++) {
      if (parents[j] < 0 || parents[j] >= nv) return "parent out of range";
      if (var_pos(family[v], parents[j]) < 0) return "family clique does not hold a parent";
    }
    const int np = prior_off[v + 1] - prior_off[v];
    if (np != (nparents(v) == 0 ? card[v] : 0)) return "prior array does not match parentless variables";
  }
  csize.resize(nc);
  for (int c = 0; c < nc; c++) {
    int64_t n = 1;
    for (int k = 0; k < clique_dim(c); k++) {
      const int v = clique_vars(c)[k];
      if (v < 0 || v >= nv) return "clique variable out of range";
      n *= card[v];
      if (n > INT32_MAX) return "clique table exceeds 2^31 entries (the reference's int index limit)";
    }
    if (n != toff[c + 1] - toff[c]) return "clique table size does not match its variables";
    csize[c] = (int)n;
  }
  if (ns != nc - 1) return "join tree must be connected: expected n_cliques-1 sepsets";
  ssize.resize(ns);
  for (int s = 0; s < ns; s++) {
    int n = 1;
    for (int k = 0; k < sepset_dim(s); k++) {
      const int v = sepset_vars(s)[k];
      if (var_pos(scl[2 * s], v) < 0 || var_pos(scl[2 * s + 1], v) < 0)
        return "sepset variable missing from a neighbour clique";
      n *= card[v];
    }
    ssize[s] = n;
  }
  S = 1;
  for (int i = 0; i < nif; i++) {
    if (card[outg[i]] != card[prev[i]]) return "interface variable pair with different cardinality";
    S *= card[outg[i]];
  }
  if (nif > 0) {
    if (in_clique < 0 || in_clique >= nc || out_clique < 0 || out_clique >= nc)
      return "in/out clique missing";
    for (int i = 0; i < nif; i++)
      if (var_pos(in_clique, prev[i]) < 0 || var_pos(out_clique, outg[i]) < 0)
        return "interface clique does not hold the interface";
  }

  // ---- schedule: DFS from cliques[0] in adjacency-list order ----
  std::vector<char> mark(nc, 0);
  sep_slot.assign(ns, 0);
  msg_total = 0; msg_max = 1;
  for (int s = 0; s < ns; s++) { sep_slot[s] = msg_total; msg_total += ssize[s]; msg_max = std::max(msg_max, ssize[s]); }
  auto make_msg = [&](int src, int s, int dst) {
    Msg m;
    m.src = src; m.dst = dst; m.sepset = s;
    std::vector<int> sv(sepset_vars(s), sepset_vars(s) + sepset_dim(s));
    m.proj_src = add_proj(src, sv);
    m.proj_dst = add_proj(dst, sv);
    m.slot = sep_slot[s];
    m.size = ssize[s];
    return m;
  };
  std::vector<int> parent_of(nc, -1), parent_sep(nc, -1);
  std::function<void(int)> dfs_collect = [&](int c) {
    mark[c] = 1;
    for (int l = adjoff[c]; l < adjoff[c + 1]; l++) {
      const int s = adj[l];
      for (int side = 0; side < 2; side++) {
        const int nb = scl[2 * s + side];
        if (!mark[nb]) {
          parent_of[nb] = c; parent_sep[nb] = s;
          dfs_collect(nb);
          collect.push_back(make_msg(nb, s, c));
        }
      }
    }
  };
  dfs_collect(0);
  for (int c = 0; c < nc; c++)
    if (!mark[c]) return "join tree is not connected";
  std::fill(mark.begin(), mark.end(), 0);
  std::function<void(int)> dfs_distribute = [&](int c) {
    mark[c] = 1;
    std::vector<int> kids;
    for (int l = adjoff[c]; l < adjoff[c + 1]; l++) {
      const int s = adj[l];
      const int nb = !mark[scl[2 * s]] ? scl[2 * s] : (!mark[scl[2 * s + 1]] ? scl[2 * s + 1] : -1);
      if (nb < 0 || parent_of[nb] != c || std::find(kids.begin(), kids.end(), nb) != kids.end()) continue;
      distribute.push_back(make_msg(c, s, nb));
      kids.push_back(nb);
    }
    for (int nb : kids) dfs_distribute(nb);
  };
  dfs_distribute(0);
  if ((int)collect.size() != nc - 1 || (int)distribute.size() != nc - 1) return "schedule construction failed";
  if (nif > 0) {  // messages on the root -> out_clique path, in top-down order
    std::vector<int> chain;
    for (int c = out_clique; c != 0; c = parent_of[c]) chain.push_back(c);
    std::reverse(chain.begin(), chain.end());
    for (int c : chain) path_to_out.push_back(make_msg(parent_of[c], parent_sep[c], c));
  }

  // ---- projections used by the slice loops ----
  if (nif > 0) {
    proj_in = add_proj(in_clique, prev);
    proj_out = add_proj(out_clique, outg);
  }
  proj_var.resize(nv); proj_fam.resize(nv); coff.assign(nv + 1, 0);
  fam_max = 1; card_max = 1;
  for (int v = 0; v < nv; v++) {
    proj_var[v] = add_proj(family[v], std::vector<int>{v});
    std::vector<int> fv{v};
    for (int j = poff[v]; j < poff[v + 1]; j++) fv.push_back(parents[j]);
    proj_fam[v] = add_proj(family[v], fv);
    coff[v + 1] = coff[v] + projs[proj_fam[v]].m;
    fam_max = std::max(fam_max, projs[proj_fam[v]].m);
    card_max = std::max(card_max, card[v]);
    if (nparents(v) == 0) prior_vars.push_back(v);
  }

  // ---- chain form? ----
  chain_ok = false;
  do {
    if (nif == 0) { chain_why = "no time-slice interface"; break; }
    if (in_clique != out_clique) { chain_why = "in_clique != out_clique"; break; }
    const int c0 = in_clique;
    if (clique_dim(c0) != 2 * nif) { chain_why = "interface clique holds extra variables"; break; }
    std::set<int> pv(prev.begin(), prev.end()), ov(outg.begin(), outg.end());
    bool bad = false;
    for (int s = 0; s < ns && !bad; s++) {
      const int a = scl[2 * s], b = scl[2 * s + 1];
      if (a != c0 && b != c0) { chain_why = "a clique is not adjacent to the interface clique"; bad = true; break; }
      for (int k = 0; k < sepset_dim(s); k++)
        if (!ov.count(sepset_vars(s)[k])) { chain_why = "a sepset holds a previous-slice variable"; bad = true; break; }
    }
    if (bad) break;
    for (int c = 0; c < nc; c++) {
      if (c == c0) continue;
      leaves.push_back(c);
      for (int s = 0; s < ns; s++)
        if ((scl[2 * s] == c && scl[2 * s + 1] == c0) || (scl[2 * s] == c0 && scl[2 * s + 1] == c))
          leaf_sepset.push_back(s);
    }
    if (leaf_sepset.size() != leaves.size()) { chain_why = "leaf bookkeeping"; leaves.clear(); leaf_sepset.clear(); break; }
    chain_ok = true;
  } while (0);
  return "";
}

}  // namespace nipgpu
