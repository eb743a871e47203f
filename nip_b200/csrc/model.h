// model.h — host-side "model compiler" of the device backend.
//
// Turns the flat description (include/nipgpu.h) into what the kernels consume:
//   * projections: for a clique table and an ordered list of its variables, the
//     precomputed index arrays base[m] / off[R] with
//         entry(j, r) = base[j] + off[r]
//     enumerating, for every destination entry j, the R clique entries that
//     project onto it.  They replace the per-entry nip_inverse_mapping /
//     nip_choose_potential_indices / nip_get_potential_pointer loops
//     (src/nippotential.c:251-264, 72-81, 58-68) and the per-call nip_mapper
//     (src/nipvariable.c:560-589);
//   * the message schedule: the join tree rooted at cliques[0] as in
//     make_consistent (src/nip.c:1600-1617), collect = post-order, distribute =
//     pre-order (src/nipjointree.c:580-673);
//   * the chain form (interface clique + leaf cliques) when the model has it.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "nipgpu.h"

namespace nipgpu {

struct Proj {
  int clique = -1;
  std::vector<int> vars;   // destination variables, dim 0 first
  int m = 1;               // destination size
  int R = 1;               // clique entries per destination entry
  int lanes = 1;           // threads cooperating on one destination entry (pow2 <= 32)
  std::vector<int> base, off;
  int base_pos = 0, off_pos = 0;  // positions inside the device int pool
  // the same map in table order, for kernels that stream the table linearly:
  // destination index of entry e = jhi[e / F] + jlo[e % F], F = product of leading dimensions
  int F = 1;
  std::vector<int> jlo, jhi;
  int jlo_pos = 0, jhi_pos = 0;
};

struct Msg {
  int src, dst, sepset;
  int proj_src, proj_dst;  // projections of src / dst clique onto the sepset
  int slot;                // where the collect-phase message of this sepset is kept
  int size;
};

// Host copy of the description plus everything derived from its structure.
struct HostModel {
  // ---- verbatim from the description ----
  int nv = 0, nc = 0, ns = 0, nif = 0, in_clique = -1, out_clique = -1;
  std::vector<int> card, flags, poff, parents, family, prior_off;
  std::vector<double> prior;
  std::vector<int> cvoff, cvars;
  std::vector<int64_t> toff;
  std::vector<double> tables;
  std::vector<int> scl, svoff, svars, adjoff, adj, outg, prev;

  // ---- derived ----
  std::vector<int> csize;             // entries per clique
  std::vector<int> ssize;             // entries per sepset
  int S = 1;                          // interface size |I|
  std::vector<Proj> projs;
  std::vector<Msg> collect, distribute, path_to_out;  // path: root -> out_clique subset of distribute
  std::vector<int> sep_slot;          // per sepset: offset inside the message area
  int msg_total = 0, msg_max = 1;
  int proj_in = -1, proj_out = -1;    // in_clique -> I_{t-1}, out_clique -> I_t
  std::vector<int> proj_var;          // per variable: family clique -> {v}
  std::vector<int> proj_fam;          // per variable: family clique -> (v, parents...)
  std::vector<int64_t> coff;          // per variable: offset of its family count table
  int fam_max = 1, card_max = 1;
  std::vector<int> prior_vars;        // parentless variables (model->independent[])

  // chain form (engine 2); valid when chain_ok
  bool chain_ok = false;
  std::string chain_why;              // why not, for diagnostics
  std::vector<int> leaves;            // leaf cliques (all cliques except in_clique)
  std::vector<int> leaf_sepset;       // sepset id connecting leaf -> interface clique

  int clique_dim(int c) const { return cvoff[c + 1] - cvoff[c]; }
  const int* clique_vars(int c) const { return cvars.data() + cvoff[c]; }
  int sepset_dim(int s) const { return svoff[s + 1] - svoff[s]; }
  const int* sepset_vars(int s) const { return svars.data() + svoff[s]; }
  int nparents(int v) const { return poff[v + 1] - poff[v]; }
  int var_pos(int c, int v) const;

  // returns "" on success, else a message
  std::string load(const nipgpu_model_desc* d);
  int add_proj(int clique, const std::vector<int>& vars);
};

}  // namespace nipgpu
