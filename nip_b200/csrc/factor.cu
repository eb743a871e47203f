// factor.cu — engine 3, see factor.cuh: the join tree evaluated factor by factor.
#include "factor.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <set>

namespace nipgpu {

namespace {

long long prod_card(const HostModel& hm, const std::vector<int>& vars) {
  long long n = 1;
  for (int v : vars) n *= hm.card[v];
  return n;
}

bool has_var(const std::vector<int>& vs, int v) { return std::find(vs.begin(), vs.end(), v) != vs.end(); }

// stride of variable v inside a tensor with the given variable order (0 when absent)
long long stride_in(const HostModel& hm, const std::vector<int>& vars, int v) {
  long long s = 1;
  for (int u : vars) {
    if (u == v) return s;
    s *= hm.card[u];
  }
  return 0;
}

template <class T>
int dev_up(T** dst, const std::vector<T>& src, cudaStream_t st) {
  cudaFree(*dst);
  *dst = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
  if (!src.empty())
    NIPGPU_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return NIPGPU_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// structure and factor extraction (host)
// ------------------------------------------------------------------------------------------
void fac_build(const HostModel& hm, FacEngine& fe, double rel_tol) {
  fe = FacEngine();
  const int nc = hm.nc, nv = hm.nv;
  fe.root = hm.nif > 0 ? hm.out_clique : 0;
  fe.parent.assign(nc, -1);
  fe.psep.assign(nc, -1);
  fe.children.assign(nc, {});
  std::vector<char> seen(nc, 0);
  std::vector<int> stack{fe.root};
  seen[fe.root] = 1;
  while (!stack.empty()) {   // pre-order
    const int c = stack.back();
    stack.pop_back();
    fe.preorder.push_back(c);
    for (int l = hm.adjoff[c]; l < hm.adjoff[c + 1]; l++) {
      const int s = hm.adj[l];
      const int nb = hm.scl[2 * s] == c ? hm.scl[2 * s + 1] : hm.scl[2 * s];
      if (seen[nb]) continue;
      seen[nb] = 1;
      fe.parent[nb] = c;
      fe.psep[nb] = s;
      fe.children[c].push_back(nb);
      stack.push_back(nb);
    }
  }
  if ((int)fe.preorder.size() != nc) { fe.why = "join tree is not connected"; return; }

  fe.cpt_tensor.assign(nv, -1);
  fe.prior_tensor.assign(nv, -1);
  fe.kappa_tensor.assign(nc, -1);
  fe.local.assign(nc, {});
  auto add_model = [&](const std::vector<int>& vars) {
    FacTensor t;
    t.vars = vars;
    t.kind = FT_MODEL;
    t.size = prod_card(hm, vars);
    t.off = fe.fac_total;
    fe.fac_total += (t.size + 15) / 16 * 16;
    fe.tensors.push_back(t);
    return (int)fe.tensors.size() - 1;
  };
  for (int v = 0; v < nv; v++) {
    if (hm.nparents(v) > 0) {
      std::vector<int> fv{v};
      for (int j = hm.poff[v]; j < hm.poff[v + 1]; j++) fv.push_back(hm.parents[j]);
      std::set<int> uniq(fv.begin(), fv.end());
      if (uniq.size() != fv.size()) { fe.why = "a family lists a variable twice"; return; }
      fe.cpt_tensor[v] = add_model(fv);
      fe.local[hm.family[v]].push_back(FacOpRef{fe.cpt_tensor[v], false});
    } else if (!(hm.flags[v] & NIPGPU_IF_OLD_OUTGOING)) {
      fe.prior_tensor[v] = add_model({v});
      fe.local[hm.family[v]].push_back(FacOpRef{fe.prior_tensor[v], false});
    }
  }
  for (int c = 0; c < nc; c++) fe.kappa_tensor[c] = add_model({});
  if (hm.nif > 0) fe.a0_tensor = add_model(hm.prev);
  fe.n_model_tensors = (int)fe.tensors.size();
  fe.h_fac.assign((size_t)fe.fac_total, 1.0);
  if (!fac_extract(hm, fe, hm.tables.data(), rel_tol)) return;
  for (int c = 0; c < nc; c++)
    if (fe.h_fac[(size_t)fe.tensors[fe.kappa_tensor[c]].off] != 1.0)
      fe.local[c].push_back(FacOpRef{fe.kappa_tensor[c], false});
  fe.ok = true;
}

bool fac_extract(const HostModel& hm, FacEngine& fe, const double* tables, double rel_tol) {
  const int nc = hm.nc;
  for (int c = 0; c < nc; c++) {
    const double* T = tables + hm.toff[c];
    const long long n = hm.csize[c];
    const int nd = hm.clique_dim(c);
    const int* cv = hm.clique_vars(c);
    std::vector<long long> cstride(nd);
    {
      long long s = 1;
      for (int k = 0; k < nd; k++) { cstride[k] = s; s *= hm.card[cv[k]]; }
    }
    std::vector<int> fams;   // variables whose CPT lives here
    for (int v = 0; v < hm.nv; v++)
      if (hm.family[v] == c && fe.cpt_tensor[v] >= 0) fams.push_back(v);
    // reference entry: the largest one
    long long xs = 0;
    for (long long e = 1; e < n; e++)
      if (T[e] > T[xs]) xs = e;
    if (!(T[xs] > 0)) { fe.why = "a clique table has no positive entry"; return false; }
    std::vector<int> xdig(nd);
    {
      long long rem = xs;
      for (int k = 0; k < nd; k++) { xdig[k] = (int)(rem % hm.card[cv[k]]); rem /= hm.card[cv[k]]; }
    }
    const int nf = (int)fams.size();
    // per factor: for each clique dimension its stride inside the factor (0: not a family member)
    std::vector<std::vector<long long>> fstride(nf, std::vector<long long>(nd, 0));
    std::vector<double*> fval(nf);
    for (int i = 0; i < nf; i++) {
      const FacTensor& ft = fe.tensors[fe.cpt_tensor[fams[i]]];
      for (int k = 0; k < nd; k++) fstride[i][k] = stride_in(hm, ft.vars, cv[k]);
      fval[i] = fe.h_fac.data() + ft.off;
    }
    auto fidx = [&](int i, const std::vector<int>& dig) {
      long long j = 0;
      for (int k = 0; k < nd; k++) j += dig[k] * fstride[i][k];
      return j;
    };
    // f_i(x_Fi) = T(x_Fi, x*) / prod_{j<i} f_j(...)   (0/0 -> 0)
    for (int i = 0; i < nf; i++) {
      const FacTensor& ft = fe.tensors[fe.cpt_tensor[fams[i]]];
      std::vector<int> dig = xdig;
      std::vector<int> pos(ft.vars.size());
      for (size_t a = 0; a < ft.vars.size(); a++) pos[a] = hm.var_pos(c, ft.vars[a]);
      for (long long j = 0; j < ft.size; j++) {
        long long rem = j;
        for (size_t a = 0; a < ft.vars.size(); a++) {
          dig[pos[a]] = (int)(rem % hm.card[ft.vars[a]]);
          rem /= hm.card[ft.vars[a]];
        }
        long long e = 0;
        for (int k = 0; k < nd; k++) e += dig[k] * cstride[k];
        double val = T[e];
        for (int p = 0; p < i && val != 0; p++) {
          const double d = fval[p][fidx(p, dig)];
          val = d != 0 ? val / d : 0.0;
        }
        fval[i][j] = val;
      }
    }
    double kappa = T[xs];
    for (int i = 0; i < nf; i++) {
      const double d = fval[i][fidx(i, xdig)];
      kappa = d != 0 ? kappa / d : 0.0;
    }
    fe.h_fac[(size_t)fe.tensors[fe.kappa_tensor[c]].off] = kappa;
    // verification over the whole table (odometer, no divisions)
    std::vector<int> dig(nd, 0);
    std::vector<long long> fi(nf, 0);
    for (long long e = 0; e < n; e++) {
      double p = kappa;
      for (int i = 0; i < nf; i++) p *= fval[i][fi[i]];
      const double t = T[e];
      if (!(std::fabs(p - t) <= rel_tol * std::fabs(t))) {
        fe.why = "clique " + std::to_string(c) + " is not the product of its families' tables";
        return false;
      }
      for (int k = 0; k < nd; k++) {   // next entry
        dig[k]++;
        for (int i = 0; i < nf; i++) fi[i] += fstride[i][k];
        if (dig[k] < hm.card[cv[k]]) break;
        for (int i = 0; i < nf; i++) fi[i] -= fstride[i][k] * hm.card[cv[k]];
        dig[k] = 0;
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
namespace {

struct PriorJob {
  int n;                      // unary prior tensors
  long long dst[32];
  int src[32], card[32], flag[32];
};
__global__ void k_fac_priors(PriorJob J, const double* prior, const int* flags, double* fac) {
  for (int k = blockIdx.x; k < J.n; k += gridDim.x)
    for (int i = threadIdx.x; i < J.card[k]; i += blockDim.x)
      fac[J.dst[k] + i] = flags[J.flag[k]] ? prior[J.src[k] + i] : 1.0;
}

struct A0Job {
  int nif, S;
  int card[24], src[24], flag[24];   // src < 0: the interface variable has parents (no prior)
};
__global__ void k_fac_a0(A0Job J, const double* prior, const int* flags, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J.S) return;
  int rem = j;
  double p = 1.0;
  for (int k = 0; k < J.nif; k++) {
    const int d = rem % J.card[k];
    rem /= J.card[k];
    if (J.src[k] >= 0 && flags[J.flag[k]]) p *= prior[J.src[k] + d];
  }
  out[j] = p;
}

__global__ void k_fac_fill(double* a, long long n, double v) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// what a kernel needs to find the data of "its" sequence
struct SlotCtx {
  double* slots;
  long long slot_stride;
  const double* fac;
  const int* pool;
  const int* obs;
  int n_obs;
  const long long* row_off;
  const int* order;    // sorted position -> series
  int wave0, t;
  long long saved_off; // where this slice's saved tensors start inside the slot area
};

__device__ __forceinline__ double fac_inv(double x) { return x != 0 ? 1.0 / x : 0.0; }

// result(o) [chunk] = prod_{epilogue operands}(o) * sum_{r in chunk} prod_k operand_k(o, r)
// for the TJ results of a thread at once: `NSH` in-loop operands do not depend on the tile
// variable (one load per term of the tile), `NV` do (TJ loads).
// EV: some in-loop operand is an evidence vector (a compare instead of a load); the kernels
// without such operands do not carry the test through their inner loop.
template <int NSH, int NV, int TJ, bool EV>
__global__ void __launch_bounds__(128) k_fac_contract(FacStepDev s, SlotCtx X) {
  constexpr int NR = NSH + NV;
  extern __shared__ int s_roff[];   // [NR][cpc * Rc]
  const int slot = blockIdx.z;
  const int span = s.cpc * s.Rc;                       // summed indices covered by this CTA
  const int rb0 = blockIdx.y * span, rb1 = min(s.R, rb0 + span);
  for (int k = 0; k < NR; k++)
    for (int i = threadIdx.x; i < rb1 - rb0; i += blockDim.x) s_roff[k * span + i] = X.pool[s.opR[k].roff + rb0 + i];
  __syncthreads();
  int o, chunk;
  if (s.cpc > 1) { o = threadIdx.x % s.othr; chunk = blockIdx.y * s.cpc + threadIdx.x / s.othr; }
  else { o = blockIdx.x * blockDim.x + threadIdx.x; chunk = blockIdx.y; }
  if (o >= s.n_thr || chunk >= s.n_chunks || (s.cpc > 1 && threadIdx.x / s.othr >= s.cpc)) return;
  const int r0 = chunk * s.Rc, r1 = min(s.R, r0 + s.Rc), nr = r1 - r0;
  const int* roff = s_roff + (r0 - rb0);
  const int oh = o / s.F, ol = o - oh * s.F;
  double* area = X.slots + (long long)slot * X.slot_stride;
  const long long row = X.row_off[X.order[X.wave0 + slot]] + X.t;
  const double* ptr[NR > 0 ? NR : 1];
  int ev[NR > 0 ? NR : 1], ts[NR > 0 ? NR : 1][TJ];
#pragma unroll
  for (int k = 0; k < NR; k++) {
    const FacOpDev& op = s.opR[k];
    const int ob = X.pool[op.ohi + oh] + X.pool[op.olo + ol];
#pragma unroll
    for (int j = 0; j < TJ; j++) ts[k][j] = op.toff[j];
    if (op.kind == FT_EVID) {
      const int obs = X.obs[row * X.n_obs + op.col];
      ev[k] = obs < 0 ? INT_MIN : obs - ob;   // the operand is 1 where its offset == ev (everywhere when missing)
      ptr[k] = nullptr;
    } else {
      ev[k] = 0;
      ptr[k] = (op.kind == FT_MODEL ? X.fac : area + (op.kind == FT_SAVED ? X.saved_off : 0)) + op.off + ob;
    }
  }
  double acc[TJ];
#pragma unroll
  for (int j = 0; j < TJ; j++) acc[j] = NR == 0 ? 1.0 : 0.0;
  if (NR > 0) {
    auto value = [&](int k, int off) -> double {
      if constexpr (EV) return ptr[k] ? ptr[k][off] : ((ev[k] == INT_MIN || off == ev[k]) ? 1.0 : 0.0);
      else return ptr[k][off];
    };
#pragma unroll 4
    for (int r = 0; r < nr; r++) {
      double sh = 1.0;
#pragma unroll
      for (int k = 0; k < NSH; k++) {
        const double v = value(k, roff[k * span + r]);
        sh = k == 0 ? v : sh * v;
      }
#pragma unroll
      for (int j = 0; j < TJ; j++) {
        double term = sh;
#pragma unroll
        for (int k = NSH; k < NR; k++) {
          const double v = value(k, roff[k * span + r] + ts[k][j]);
          term = (NSH == 0 && k == NSH) ? v : term * v;
        }
        acc[j] += term;
      }
    }
  }
  for (int k = 0; k < s.nO; k++) {
    const FacOpDev& op = s.opO[k];
    const int ob = X.pool[op.ohi + oh] + X.pool[op.olo + ol];
    int obs = 0;
    const double* q = nullptr;
    if (op.kind == FT_EVID) obs = X.obs[row * X.n_obs + op.col];
    else q = (op.kind == FT_MODEL ? X.fac : area + (op.kind == FT_SAVED ? X.saved_off : 0)) + op.off;
#pragma unroll
    for (int j = 0; j < TJ; j++) {
      const int off = ob + op.toff[j];
      double v;
      if (op.kind == FT_EVID) v = (obs < 0 || obs == off) ? 1.0 : 0.0;
      else {
        v = q[off];
        if (op.inv) v = fac_inv(v);
      }
      acc[j] *= v;
    }
  }
  // the result (or this chunk's partial sums, laid out like the result)
  const int oo = X.pool[s.out.ohi + oh] + X.pool[s.out.olo + ol];
  double* dst = area + (s.out.kind == FT_SAVED ? X.saved_off : 0) + s.out.off +
                (s.n_chunks > 1 ? (long long)chunk * s.n_out : 0);
#pragma unroll
  for (int j = 0; j < TJ; j++) dst[oo + s.out.toff[j]] = acc[j];
}

// The same contraction with a 2-D register tile: TJ x TK results per thread along two output
// variables such that no in-loop operand holds both.  Per term of the summed index a thread
// loads TJ values of the operands that hold the first variable and TK of those that hold the
// second, and forms their outer product: (TJ + TK) loads for TJ * TK multiply-adds.
template <int NSH, int NV, int NV2, int TJ, int TK>
__global__ void __launch_bounds__(128) k_fac_contract2(FacStepDev s, SlotCtx X) {
  constexpr int NR = NSH + NV + NV2;
  extern __shared__ int s_roff[];   // [NR][cpc * Rc]
  const int slot = blockIdx.z;
  const int span = s.cpc * s.Rc;
  const int rb0 = blockIdx.y * span, rb1 = min(s.R, rb0 + span);
  for (int k = 0; k < NR; k++)
    for (int i = threadIdx.x; i < rb1 - rb0; i += blockDim.x) s_roff[k * span + i] = X.pool[s.opR[k].roff + rb0 + i];
  __syncthreads();
  int o, chunk;
  if (s.cpc > 1) { o = threadIdx.x % s.othr; chunk = blockIdx.y * s.cpc + threadIdx.x / s.othr; }
  else { o = blockIdx.x * blockDim.x + threadIdx.x; chunk = blockIdx.y; }
  if (o >= s.n_thr || chunk >= s.n_chunks || (s.cpc > 1 && threadIdx.x / s.othr >= s.cpc)) return;
  const int r0 = chunk * s.Rc, r1 = min(s.R, r0 + s.Rc), nr = r1 - r0;
  const int* roff = s_roff + (r0 - rb0);
  const int oh = o / s.F, ol = o - oh * s.F;
  double* area = X.slots + (long long)slot * X.slot_stride;
  const long long row = X.row_off[X.order[X.wave0 + slot]] + X.t;
  const double* ptr[NR];
  int ev[NR], ts[NR][4];
#pragma unroll
  for (int k = 0; k < NR; k++) {
    const FacOpDev& op = s.opR[k];
    const int ob = X.pool[op.ohi + oh] + X.pool[op.olo + ol];
#pragma unroll
    for (int j = 0; j < 4; j++) ts[k][j] = k < NSH + NV ? op.toff[j] : op.toff2[j];
    if (op.kind == FT_EVID) {
      const int obs = X.obs[row * X.n_obs + op.col];
      ev[k] = obs < 0 ? INT_MIN : obs - ob;
      ptr[k] = nullptr;
    } else {
      ev[k] = 0;
      ptr[k] = (op.kind == FT_MODEL ? X.fac : area + (op.kind == FT_SAVED ? X.saved_off : 0)) + op.off + ob;
    }
  }
  auto value = [&](int k, int off) -> double { return ptr[k][off]; };   // the planner keeps evidence operands out
  double acc[TJ][TK];
#pragma unroll
  for (int j = 0; j < TJ; j++)
#pragma unroll
    for (int q = 0; q < TK; q++) acc[j][q] = 0.0;
#pragma unroll 2
  for (int r = 0; r < nr; r++) {
    double a[TJ], b[TK];
#pragma unroll
    for (int j = 0; j < TJ; j++) {
      a[j] = value(NSH, roff[NSH * span + r] + ts[NSH][j]);
#pragma unroll
      for (int k = NSH + 1; k < NSH + NV; k++) a[j] *= value(k, roff[k * span + r] + ts[k][j]);
    }
#pragma unroll
    for (int q = 0; q < TK; q++) {
      b[q] = value(NSH + NV, roff[(NSH + NV) * span + r] + ts[NSH + NV][q]);
#pragma unroll
      for (int k = NSH + NV + 1; k < NR; k++) b[q] *= value(k, roff[k * span + r] + ts[k][q]);
    }
    if (NSH > 0) {
      double sh = value(0, roff[r]);
#pragma unroll
      for (int k = 1; k < NSH; k++) sh *= value(k, roff[k * span + r]);
#pragma unroll
      for (int q = 0; q < TK; q++) b[q] *= sh;
    }
#pragma unroll
    for (int j = 0; j < TJ; j++)
#pragma unroll
      for (int q = 0; q < TK; q++) acc[j][q] = fma(a[j], b[q], acc[j][q]);
  }
  for (int k = 0; k < s.nO; k++) {
    const FacOpDev& op = s.opO[k];
    const int ob = X.pool[op.ohi + oh] + X.pool[op.olo + ol];
    int obs = 0;
    const double* qp = nullptr;
    if (op.kind == FT_EVID) obs = X.obs[row * X.n_obs + op.col];
    else qp = (op.kind == FT_MODEL ? X.fac : area + (op.kind == FT_SAVED ? X.saved_off : 0)) + op.off;
#pragma unroll
    for (int j = 0; j < TJ; j++)
#pragma unroll
      for (int q = 0; q < TK; q++) {
        const int off = ob + op.toff[j] + op.toff2[q];
        double v;
        if (op.kind == FT_EVID) v = (obs < 0 || obs == off) ? 1.0 : 0.0;
        else {
          v = qp[off];
          if (op.inv) v = fac_inv(v);
        }
        acc[j][q] *= v;
      }
  }
  const int oo = X.pool[s.out.ohi + oh] + X.pool[s.out.olo + ol];
  double* dst = area + (s.out.kind == FT_SAVED ? X.saved_off : 0) + s.out.off +
                (s.n_chunks > 1 ? (long long)chunk * s.n_out : 0);
#pragma unroll
  for (int j = 0; j < TJ; j++)
#pragma unroll
    for (int q = 0; q < TK; q++) dst[oo + s.out.toff[j] + s.out.toff2[q]] = acc[j][q];
}

// out[o] = sum over chunks of partial[chunk][o], fixed order
__global__ void k_fac_reduce(SlotCtx X, long long part_off, long long out_off, int n_out, int n_chunks) {
  double* area = X.slots + (long long)blockIdx.y * X.slot_stride;   // out_off already holds the saved shift
  const double* part = area + part_off;
  if (n_chunks >= 32) {   // a warp per output
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= n_out) return;
    double s = 0;
    for (int c = lane; c < n_chunks; c += 32) s += part[(long long)c * n_out + o];
    s = warp_sum(s);
    if (lane == 0) area[out_off + o] = s;
  } else {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    double s = 0;
    for (int c = 0; c < n_chunks; c++) s += part[(long long)c * n_out + o];
    area[out_off + o] = s;
  }
}

// start of a slice: alpha_in <- alpha_{t-1} (or the prior product on first slices); backward:
// beta <- 1 on the last slice of a sequence
// (the forward rows live in a store of the wave, [slot][t_rows][S]: a sequence needs them from its
// own forward pass to its own backward pass only, so HBM holds W x T of them, not one per data row)
__global__ void k_fac_prepare(SlotCtx X, const int* len_sorted, const double* alpha, int t_rows, int S,
                              long long o_alpha_in, long long o_beta, long long a0_off,
                              int set_beta /*0 no, 1 where last, 2 all*/, double* ll_run, int* bad_run, int reset_ll) {
  const int slot = blockIdx.y;
  double* area = X.slots + (long long)slot * X.slot_stride;
  const int p = X.wave0 + slot;
  const double* src = X.t == 0 ? X.fac + a0_off : alpha + ((long long)slot * t_rows + X.t - 1) * S;
  const bool beta1 = set_beta == 2 || (set_beta == 1 && len_sorted[p] - 1 == X.t);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    area[o_alpha_in + i] = src[i];
    if (beta1) area[o_beta + i] = 1.0;
  }
  if (reset_ll && blockIdx.x == 0 && threadIdx.x == 0) { ll_run[slot] = 0.0; bad_run[slot] = 0; }
}

// end of a forward slice: masses, log-likelihood bookkeeping (src/nip.c:1458-1474, 1827-1831),
// alpha_t = normalise(alpha_new) into the row store
__global__ void __launch_bounds__(1024) k_fac_settle_fwd(SlotCtx X, const int* len_sorted, const int* marked, int S,
                                                         long long o_alpha_in, long long o_alpha_new,
                                                         const double* R1, const double* m1_0, double* alpha,
                                                         int t_rows, int want_ll, int nif, double* ll_run,
                                                         int* bad_run, double* ll_out, int* status_out) {
  __shared__ double red[40];
  const int slot = blockIdx.x;
  double* area = X.slots + (long long)slot * X.slot_stride;
  const int p = X.wave0 + slot, seq = X.order[p];
  const long long row0 = X.row_off[seq];
  const double* an = area + o_alpha_new;
  double s = 0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += an[i];
  double m2 = block_sum(s, red);
  const double inv = m2 != 0 ? 1.0 / m2 : 1.0;      // zero sum: untouched (nip_normalise_array)
  double* arow = alpha + ((long long)slot * t_rows + X.t) * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) arow[i] = an[i] * inv;
  if (want_ll) {
    double m1;
    if (X.t == 0) m1 = *m1_0;
    else if (nif == 0) m1 = R1[0];
    else {
      double d = 0;
      for (int i = threadIdx.x; i < S; i += blockDim.x) d += area[o_alpha_in + i] * R1[i];
      m1 = block_sum(d, red);
    }
    if (threadIdx.x == 0) {
      int entered = 0;
      const int* obs = X.obs + (row0 + X.t) * X.n_obs;
      for (int k = 0; k < X.n_obs; k++) entered += marked[k] && obs[k] >= 0;
      if (entered == 0) m2 = m1;   // an evidence-free slice has m2 == m1 by definition
      double ll = ll_run[slot];
      int bad = bad_run[slot];
      if (m1 > 0 && m2 > 0) ll += log(m2) - log(m1);
      if (m2 == 0) ll = -DBL_MAX;
      if (m1 <= 0 || m2 <= 0 || ll > 0) bad = 1;
      ll_run[slot] = ll;
      bad_run[slot] = bad;
      if (X.t == len_sorted[p] - 1) {
        if (ll_out) ll_out[seq] = ll;
        if (status_out) status_out[seq] = bad;
      }
    }
  } else if (threadIdx.x == 0 && X.t == len_sorted[p] - 1) {
    if (ll_out) ll_out[seq] = 0.0;
    if (status_out) status_out[seq] = 0;
  }
}

// beta_{t-1} = b / sum(b)   (any positive scale gives the same posteriors; this keeps it O(1))
__global__ void __launch_bounds__(1024) k_fac_beta_norm(SlotCtx X, int S, long long o_bprev, long long o_beta) {
  __shared__ double red[40];
  double* area = X.slots + (long long)blockIdx.x * X.slot_stride;
  double s = 0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += area[o_bprev + i];
  s = block_sum(s, red);
  const double inv = s != 0 ? 1.0 / s : 1.0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) area[o_beta + i] = area[o_bprev + i] * inv;
}

// expected counts: acc[slot][dst + i] += v[i] / sum(v)   (src/nip.c:1925-1967; zero mass: skipped)
__global__ void __launch_bounds__(1024) k_fac_count(SlotCtx X, long long src_off, int n, double* acc,
                                                    long long acc_stride, long long dst) {
  __shared__ double red[40];
  const double* v = X.slots + (long long)blockIdx.x * X.slot_stride + src_off;
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (s == 0) return;
  double* a = acc + (long long)blockIdx.x * acc_stride + dst;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] += v[i] / s;
}

// posterior of a queried variable into the caller's row
__global__ void __launch_bounds__(256) k_fac_query(SlotCtx X, long long src_off, int n, double* post, int post_row,
                                                   long long dst) {
  __shared__ double red[40];
  const double* v = X.slots + (long long)blockIdx.x * X.slot_stride + src_off;
  const long long row = X.row_off[X.order[X.wave0 + blockIdx.x]] + X.t;
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  const double inv = s != 0 ? 1.0 / s : 1.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) post[row * post_row + dst + i] = v[i] * inv;
}

}  // namespace

int fac_refresh(const HostModel& hm, FacEngine& fe, const double* d_prior, const int* d_prior_flags,
                const double* d_counts, cudaStream_t st) {
  if (!fe.ok) return NIPGPU_OK;
  if (!fe.d_fac) NIPGPU_CUDA(cudaMalloc((void**)&fe.d_fac, std::max<long long>(fe.fac_total, 1) * sizeof(double)));
  if (!d_counts) {
    NIPGPU_CUDA(cudaMemcpyAsync(fe.d_fac, fe.h_fac.data(), (size_t)fe.fac_total * sizeof(double),
                                cudaMemcpyHostToDevice, st));
  } else {
    for (int v = 0; v < hm.nv; v++)
      if (fe.cpt_tensor[v] >= 0) {
        const FacTensor& t = fe.tensors[fe.cpt_tensor[v]];
        NIPGPU_CUDA(cudaMemcpyAsync(fe.d_fac + t.off, d_counts + hm.coff[v], (size_t)t.size * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
      }
    for (int c = 0; c < hm.nc; c++) {   // tables rebuilt by init_potential from all-ones: no constant left
      k_fac_fill<<<1, 32, 0, st>>>(fe.d_fac + fe.tensors[fe.kappa_tensor[c]].off, 1, 1.0);
      NIPGPU_LAUNCHED();
    }
  }
  // priors: entered only when some entry is positive (nip_enter_prior, src/nipjointree.c:917-927)
  std::vector<int> flag_of(hm.nv, -1);
  for (size_t k = 0; k < hm.prior_vars.size(); k++) flag_of[hm.prior_vars[k]] = (int)k;
  PriorJob J;
  J.n = 0;
  auto flush = [&]() -> int {
    if (J.n == 0) return NIPGPU_OK;
    k_fac_priors<<<J.n, 64, 0, st>>>(J, d_prior, d_prior_flags, fe.d_fac);
    NIPGPU_LAUNCHED();
    J.n = 0;
    return NIPGPU_OK;
  };
  for (int v = 0; v < hm.nv; v++) {
    if (fe.prior_tensor[v] < 0) continue;
    J.dst[J.n] = fe.tensors[fe.prior_tensor[v]].off;
    J.src[J.n] = hm.prior_off[v];
    J.card[J.n] = hm.card[v];
    J.flag[J.n] = flag_of[v];
    if (++J.n == 32)
      if (int e = flush()) return e;
  }
  if (int e = flush()) return e;
  if (fe.a0_tensor >= 0) {
    if (hm.nif > 24) { set_error("factor engine: more than 24 interface variables"); return NIPGPU_EUNSUPPORTED; }
    A0Job A;
    A.nif = hm.nif;
    A.S = hm.S;
    for (int k = 0; k < hm.nif; k++) {
      const int v = hm.prev[k];
      A.card[k] = hm.card[v];
      A.src[k] = hm.nparents(v) == 0 ? hm.prior_off[v] : -1;
      A.flag[k] = hm.nparents(v) == 0 ? flag_of[v] : 0;
    }
    k_fac_a0<<<(hm.S + 255) / 256, 256, 0, st>>>(A, d_prior, d_prior_flags, fe.d_fac + fe.tensors[fe.a0_tensor].off);
    NIPGPU_LAUNCHED();
  }
  return NIPGPU_OK;
}

void fac_free(FacEngine& fe) {
  cudaFree(fe.d_fac);
  cudaFree(fe.d_slots);
  cudaFree(fe.d_ll_run);
  cudaFree(fe.d_bad_run);
  cudaFree(fe.d_alpha_wave);
  for (auto& kv : fe.programs) {
    cudaFree(kv.second.d_pool);
    cudaFree(kv.second.d_marked);
  }
  fe = FacEngine();
}

// ------------------------------------------------------------------------------------------
// program compilation (host)
// ------------------------------------------------------------------------------------------
namespace {

struct Compiler {
  const HostModel& hm;
  const FacEngine& fe;
  std::vector<FacTensor> T;
  std::vector<int> pool;
  long long slot_top = 0, saved_top = 0;
  long long partial_max = 0;
  bool failed = false;
  bool saving = false;      // tensors created now go to the per-slice saved region
  // in-loop operands a contraction may have (NIPGPU_FACTOR_MAX_OPS lowers it: the tests use that to
  // exercise the fallback to engine 1)
  int max_ops = [] {
    const char* p = getenv("NIPGPU_FACTOR_MAX_OPS");
    const int v = p ? atoi(p) : 0;
    return v >= 1 && v < kFacMaxOps ? v : kFacMaxOps;
  }();
  std::map<std::vector<int>, int> relayouts;   // (tensor, first variable) -> permuted copy, valid in the current section

  Compiler(const HostModel& h, const FacEngine& f) : hm(h), fe(f), T(f.tensors.begin(), f.tensors.begin() + f.n_model_tensors) {}

  int slot_tensor(const std::vector<int>& vars) {
    FacTensor t;
    t.vars = vars;
    t.kind = saving ? FT_SAVED : FT_SLOT;
    t.size = prod_card(hm, vars);
    long long& top = saving ? saved_top : slot_top;
    t.off = top;
    top += (t.size + 15) / 16 * 16;
    T.push_back(t);
    return (int)T.size() - 1;
  }
  int evid_tensor(int v, int col) {
    FacTensor t;
    t.vars = {v};
    t.kind = FT_EVID;
    t.size = hm.card[v];
    t.col = col;
    T.push_back(t);
    return (int)T.size() - 1;
  }

  // index tables of one tensor over the thread space (variables `tv`; a tile variable's digit counts
  // blocks of `tmult[a]` states) and over the summed variables
  void tables(const FacTensor& t, const std::vector<int>& tv, const std::vector<int>& tmult, int F, long long n_thr,
              const std::vector<int>& rvars, long long R, FacOpDev& d) {
    auto fill = [&](const std::vector<int>& vars, const std::vector<int>* mult, long long count, long long unit) {
      const int pos = (int)pool.size();
      for (long long x = 0; x < count; x++) {
        long long rem = x * unit, o = 0;
        for (size_t a = 0; a < vars.size(); a++) {
          const int v = vars[a];
          const int lm = mult ? (*mult)[a] : 1;
          const int c = hm.card[v] / lm;
          o += (rem % c) * lm * stride_in(hm, t.vars, v);
          rem /= c;
        }
        pool.push_back((int)o);
      }
      return pos;
    };
    d.olo = fill(tv, &tmult, F, 1);
    d.ohi = fill(tv, &tmult, (n_thr + F - 1) / F, F);
    d.roff = fill(rvars, nullptr, R, 1);
  }

  // Adds to `prog` the contraction out(ovars) = sum over the other variables of prod(ops).
  // fixed: the result keeps the variable order of `ovars`; otherwise the order is chosen here
  // (the variable along which consecutive threads run comes first).  Returns the result tensor.
  int contract(std::vector<FacInstr>& prog, const std::vector<int>& ovars_in, bool fixed, std::vector<FacOpRef> ops,
               bool needs_history = false, bool relayout = true) {
    std::vector<int> rvars;
    for (const FacOpRef& o : ops)
      for (int v : T[o.tensor].vars)
        if (!has_var(ovars_in, v) && !has_var(rvars, v)) rvars.push_back(v);
    // A contraction that sums several variables over three or more operands: when one summed
    // variable is held by only some of them, it is summed out of THOSE first (an intermediate over
    // the variables they hold) — fewer operands per term in both halves, and halves with two
    // in-loop operands take the 2-D register tile.
    static const bool no_split = [] { const char* p = getenv("NIPGPU_FACTOR_SPLIT"); return p && p[0] == '0'; }();
    if (rvars.size() >= 2 && !no_split) {
      auto in_loop = [&](const FacTensor& t) {
        for (int v : t.vars)
          if (has_var(rvars, v)) return true;
        return false;
      };
      int n_in = 0;
      for (const FacOpRef& o : ops) n_in += in_loop(T[o.tensor]) ? 1 : 0;
      int best_r = -1;
      long long best_size = 0;
      std::vector<int> best_vars;
      if (n_in >= 3)
        for (int r : rvars) {
          std::vector<int> iv;
          int n_r = 0;
          for (const FacOpRef& o : ops) {
            const FacTensor& t = T[o.tensor];
            if (!has_var(t.vars, r)) continue;
            n_r++;
            for (int v : t.vars)
              if (v != r && !has_var(iv, v)) iv.push_back(v);
          }
          if (n_r == 0 || n_r >= n_in) continue;
          const long long size = prod_card(hm, iv);
          if (size > (1 << 21)) continue;
          if (best_r < 0 || size < best_size) { best_r = r; best_size = size; best_vars = iv; }
        }
      if (best_r >= 0) {
        std::vector<FacOpRef> first, rest;
        for (const FacOpRef& o : ops) (has_var(T[o.tensor].vars, best_r) ? first : rest).push_back(o);
        const int mid = contract(prog, best_vars, false, first, needs_history, relayout);
        rest.push_back(FacOpRef{mid, false});
        return contract(prog, ovars_in, fixed, rest, needs_history, relayout);
      }
    }
    const long long R = prod_card(hm, rvars);
    auto depends = [&](const FacTensor& t) {
      for (int v : t.vars)
        if (has_var(rvars, v)) return true;
      return false;
    };
    // ---- the variable consecutive threads run along ----
    const long long kBig = 1 << 17;     // tensors above this are not copied into another order
    int v0 = ovars_in.empty() ? -1 : ovars_in[0];
    if (ovars_in.size() > 1) {
      // the heaviest in-loop operand: among equally good directions, run along its fastest variable
      const FacTensor* heavy = nullptr;
      for (const FacOpRef& o : ops) {
        const FacTensor& t = T[o.tensor];
        if (t.kind != FT_EVID && depends(t) && (!heavy || t.size > heavy->size)) heavy = &t;
      }
      double best = -1;
      bool best_heavy = false;
      for (int v : ovars_in) {
        double sc = 0;
        for (const FacOpRef& o : ops) {
          const FacTensor& t = T[o.tensor];
          if (t.kind == FT_EVID || t.vars.empty() || t.vars[0] == v) continue;
          if (!has_var(t.vars, v)) {
            // absent: every lane reads the same entry.  Fine when the operand's fastest variable is
            // summed (a thread then walks along a line); when it is another result variable, a big
            // operand is fetched one 8-byte entry per 32-byte sector
            if (t.size > kBig && has_var(ovars_in, t.vars[0])) sc += 4.0 * t.size;
            continue;
          }
          sc += t.size > kBig ? 64.0 * t.size * (depends(t) ? (double)R : 1.0) : (double)t.size;
        }
        // a result whose order is fixed is written with a stride when the threads run along
        // another variable: one scattered store per result, nothing against the R terms behind it
        if (fixed && v != ovars_in[0]) sc += 2.0 * (double)prod_card(hm, ovars_in);
        const bool on_heavy = heavy && !heavy->vars.empty() && heavy->vars[0] == v;
        if (best < 0 || sc < best || (sc == best && on_heavy && !best_heavy)) { best = sc; v0 = v; best_heavy = on_heavy; }
      }
    }
    // operands that hold v0 but not as their fastest variable: a copy in another order, if small
    if (v0 >= 0 && relayout)
      for (FacOpRef& o : ops) {
        const FacTensor t = T[o.tensor];
        if (t.kind == FT_EVID || !has_var(t.vars, v0) || t.vars[0] == v0 || t.size > kBig || t.vars.size() < 2) continue;
        const std::vector<int> key{o.tensor, v0};
        auto it = relayouts.find(key);
        if (it == relayouts.end()) {
          std::vector<int> nv{v0};
          for (int v : t.vars)
            if (v != v0) nv.push_back(v);
          const int copy = contract(prog, nv, true, {FacOpRef{o.tensor, false}}, false, false);
          it = relayouts.emplace(key, copy).first;
        }
        o.tensor = it->second;
      }
    std::vector<int> ovars = ovars_in;
    if (!fixed && v0 >= 0) {
      ovars.clear();
      ovars.push_back(v0);
      for (int v : ovars_in)
        if (v != v0) ovars.push_back(v);
    }
    const int out = slot_tensor(ovars);
    // ---- register tiles.  A tile holds four results: along one output variable whose cardinality
    // is a multiple of 4, or over 2 x 2 states of two variables of even cardinality (binary
    // variables).  First tile: the one the in-loop operands depend on least (they are loaded once
    // per term of the tile when they do not hold its variables).
    typedef std::vector<std::pair<int, int>> Tile;   // (variable, states per thread)
    auto holds = [&](const FacTensor& t, const Tile& tl) {
      for (const auto& e : tl)
        if (has_var(t.vars, e.first)) return true;
      return false;
    };
    std::vector<Tile> candidates;
    for (int v : ovars) {
      if (v == v0 && ovars.size() > 1) continue;
      if (hm.card[v] % 4 == 0) candidates.push_back(Tile{{v, 4}});
    }
    for (size_t a = 0; a < ovars.size(); a++)
      for (size_t b = a + 1; b < ovars.size(); b++) {
        const int va = ovars[a], vb = ovars[b];
        if (va == v0 || vb == v0) continue;
        if (hm.card[va] % 2 != 0 || hm.card[vb] % 2 != 0) continue;
        if (hm.card[va] % 4 == 0 && hm.card[vb] % 4 == 0) continue;   // each of them can carry a tile alone
        candidates.push_back(Tile{{va, 2}, {vb, 2}});
      }
    Tile tile1, tile2;
    int TJ = 1, TK = 1;
    if (R > 1) {
      double best = -1;
      for (const Tile& tl : candidates) {
        double sc = 0;
        int nvar = 0;
        for (const FacOpRef& o : ops) {
          const FacTensor& t = T[o.tensor];
          if (!depends(t) || !holds(t, tl)) continue;
          sc += (double)t.size;
          nvar++;
        }
        if (nvar > kFacMaxVar) continue;
        if (best < 0 || sc < best) { best = sc; tile1 = tl; }
      }
      if (!tile1.empty()) TJ = 4;
    }
    // ---- a second tile: no in-loop operand may hold variables of both (outer product of two slices) ----
    static const bool no_2d = [] { const char* p = getenv("NIPGPU_FACTOR_TILE2"); return p && p[0] == '0'; }();
    bool evid_in_loop = false;
    for (const FacOpRef& o : ops) evid_in_loop = evid_in_loop || (T[o.tensor].kind == FT_EVID && depends(T[o.tensor]));
    if (!tile1.empty() && !no_2d && !evid_in_loop) {
      int n1 = 0, nsh = 0;
      for (const FacOpRef& o : ops) {
        const FacTensor& t = T[o.tensor];
        if (!depends(t)) continue;
        (holds(t, tile1) ? n1 : nsh)++;
      }
      double best = -1;
      for (const Tile& tl : candidates) {
        bool overlap = false;
        for (const auto& e : tl)
          for (const auto& f : tile1) overlap = overlap || e.first == f.first;
        if (overlap) continue;
        double sc = 0;
        int n2 = 0;
        bool both = false;
        for (const FacOpRef& o : ops) {
          const FacTensor& t = T[o.tensor];
          if (!depends(t) || !holds(t, tl)) continue;
          if (holds(t, tile1)) both = true;
          sc += (double)t.size;
          n2++;
        }
        // templates exist for 1..2 operands per tile and up to 2 shared ones
        if (both || n2 < 1 || n2 > 2 || n1 < 1 || n1 > 2 || nsh - n2 > 2 || nsh - n2 < 0) continue;
        if (best < 0 || sc < best) { best = sc; tile2 = tl; }
      }
      if (!tile2.empty()) TK = 4;
    }
    auto in_tile = [&](int v) {
      for (const auto& e : tile1) if (e.first == v) return e.second;
      for (const auto& e : tile2) if (e.first == v) return e.second;
      return 0;
    };
    // thread space: v0 first, then the blocks of the tile variables (threads that share the heavy
    // operands' entries sit next to each other: one trip to HBM, the rest L1/L2 hits), then the rest
    std::vector<int> tv, tmult;
    if (v0 >= 0 && !in_tile(v0)) { tv.push_back(v0); tmult.push_back(1); }
    for (const auto& e : tile1) { tv.push_back(e.first); tmult.push_back(e.second); }
    for (const auto& e : tile2) { tv.push_back(e.first); tmult.push_back(e.second); }
    for (int v : ovars)
      if (v != v0 && !in_tile(v)) { tv.push_back(v); tmult.push_back(1); }
    // offsets of a tile's four elements inside a tensor (element j: first tile variable fastest)
    auto tile_offsets = [&](const FacTensor& t, const Tile& tl, int* out4) {
      for (int j = 0; j < 4; j++) {
        int rem = j;
        long long o = 0;
        for (const auto& e : tl) {
          o += (rem % e.second) * stride_in(hm, t.vars, e.first);
          rem /= e.second;
        }
        out4[j] = tl.empty() ? 0 : (int)o;
      }
    };
    const int vt = tile1.empty() ? -1 : tile1[0].first, vt2 = tile2.empty() ? -1 : tile2[0].first;   // (trace only)
    const long long n_out = T[out].size, n_thr = n_out / (TJ * TK);
    FacInstr ins;
    ins.kind = FI_CONTRACT;
    ins.needs_history = needs_history;
    FacStepDev& s = ins.step;
    int F = 1;
    for (size_t a = 0; a < tv.size(); a++) {
      if (F >= 128) break;
      F *= hm.card[tv[a]] / tmult[a];
    }
    if (n_out > INT32_MAX || R > INT32_MAX) { failed = true; return out; }
    s.n_thr = (int)n_thr; s.n_out = (int)n_out; s.F = F; s.R = (int)R; s.TJ = TJ; s.TK = TK;
    s.nSh = s.nVar = s.nVar2 = s.nO = 0;
    auto dev = [&](const FacTensor& t, bool inv, bool dep) {
      FacOpDev d{};
      d.off = t.off; d.kind = t.kind; d.inv = inv ? 1 : 0; d.col = t.col;
      tile_offsets(t, tile1, d.toff);
      tile_offsets(t, tile2, d.toff2);
      tables(t, tv, tmult, F, n_thr, rvars, dep ? R : 0, d);
      return d;
    };
    std::vector<FacOpDev> shared, varying, varying2;
    for (const FacOpRef& o : ops) {
      const FacTensor& t = T[o.tensor];
      const bool dep = depends(t);
      const FacOpDev d = dev(t, o.inv, dep);
      if (!dep) {
        if (s.nO >= kFacMaxOps) { failed = true; return out; }
        s.opO[s.nO++] = d;
      } else {
        if (o.inv) { failed = true; return out; }
        (holds(t, tile1) ? varying : holds(t, tile2) ? varying2 : shared).push_back(d);
      }
    }
    if ((int)(shared.size() + varying.size() + varying2.size()) > max_ops || (int)varying.size() > kFacMaxVar) { failed = true; return out; }
    if (TJ > 1 && TK == 1 && (int)shared.size() > 4) { failed = true; return out; }
    for (const FacOpDev& d : shared) s.opR[s.nSh++] = d;
    for (const FacOpDev& d : varying) s.opR[s.nSh + s.nVar++] = d;
    for (const FacOpDev& d : varying2) s.opR[s.nSh + s.nVar + s.nVar2++] = d;
    s.out = dev(T[out], false, false);
    // split the summed range so that a slot has at least ~64k threads, at least 8 terms each
    const long long want = 65536;
    long long chunks = std::max<long long>(1, std::min<long long>((want + n_thr - 1) / n_thr, R / 8));
    chunks = std::min<long long>(chunks, 2048);     // the partial sums are added up by one warp per result
    chunks = std::max(chunks, (R + 1023) / 1024);
    if (s.nSh + s.nVar + s.nVar2 == 0) chunks = 1;
    s.Rc = (int)((R + chunks - 1) / chunks);
    s.n_chunks = (int)((R + s.Rc - 1) / s.Rc);
    s.cpc = 1;
    s.othr = 128;
    if (n_thr < 128 && s.n_chunks > 1) {   // few results: several chunks share a CTA
      int p2 = 1;
      while (p2 < n_thr) p2 *= 2;
      s.othr = p2;
      s.cpc = std::max(1, std::min(128 / p2, 2048 / std::max(s.Rc, 1)));
    }
    if (s.n_chunks > 1) partial_max = std::max(partial_max, (long long)s.n_chunks * n_out);
    ins.flops = (double)n_out * (double)R * std::max(1, s.nSh + s.nVar + s.nVar2);
    {
      auto vs = [&](const std::vector<int>& v) { std::string r = "("; for (int x : v) r += std::to_string(x) + " "; return r + ")"; };
      ins.note = "out" + vs(ovars) + " v0=" + std::to_string(v0) + " vt=" + std::to_string(vt) + " vt2=" + std::to_string(vt2) + " :";
      for (const FacOpRef& o : ops) ins.note += " " + std::string(T[o.tensor].kind == FT_MODEL ? "M" : T[o.tensor].kind == FT_EVID ? "E" : "S") + vs(T[o.tensor].vars);
    }
    prog.push_back(ins);
    return out;
  }
};

std::vector<int> plan_key(const FacRunArgs& a, const HostModel& hm, bool counts) {
  std::vector<int> key;
  key.push_back(a.n_obs);
  for (int k = 0; k < a.n_obs; k++) {
    const int v = (*a.obs_vars)[k];
    key.push_back((a.use_evidence ? a.use_evidence[v] != 0 : true) ? v : -1);
  }
  key.push_back(-2);
  key.push_back(counts ? 1 : 0);
  key.push_back(a.d_post ? a.n_query : 0);
  if (a.d_post)
    for (int q = 0; q < a.n_query; q++) key.push_back(a.query[q]);
  (void)hm;
  return key;
}

int compile(const HostModel& hm, FacEngine& fe, const FacRunArgs& a, bool counts, FacProgram& P, cudaStream_t st) {
  Compiler C(hm, fe);
  const int nc = hm.nc;
  const bool queries = a.d_post && a.n_query > 0;
  // ---- per-clique operand lists ----
  std::vector<std::vector<FacOpRef>> L = fe.local;
  std::vector<int> marked(std::max(a.n_obs, 1), 0);
  std::vector<char> var_seen(hm.nv, 0);
  for (int k = 0; k < a.n_obs; k++) {
    const int v = (*a.obs_vars)[k];
    if (v < 0 || v >= hm.nv) { set_error("factor engine: observed variable out of range"); return NIPGPU_EINVAL; }
    const bool on = a.use_evidence ? a.use_evidence[v] != 0 : true;
    if (!on) continue;
    marked[k] = 1;
    L[hm.family[v]].push_back(FacOpRef{C.evid_tensor(v, k), false});
    var_seen[v] = 1;
  }
  int t_alpha_in = -1, t_beta = -1;
  if (hm.nif > 0) {
    t_alpha_in = C.slot_tensor(hm.prev);
    t_beta = C.slot_tensor(hm.outg);
    P.o_alpha_in = C.T[t_alpha_in].off;
    P.o_beta = C.T[t_beta].off;
  }
  auto sep_vars = [&](int s) { return std::vector<int>(hm.sepset_vars(s), hm.sepset_vars(s) + hm.sepset_dim(s)); };
  // ---- upward messages (towards the root = out_clique), children before parents ----
  std::vector<int> up(nc, -1), down(nc, -1);
  std::vector<FacInstr> ups;
  C.saving = true;
  for (int i = nc - 1; i >= 0; i--) {
    const int c = fe.preorder[i];
    if (c == fe.root) continue;
    std::vector<FacOpRef> ops = L[c];
    if (hm.nif > 0 && c == hm.in_clique) ops.push_back(FacOpRef{t_alpha_in, false});
    for (int d : fe.children[c]) ops.push_back(FacOpRef{up[d], false});
    up[c] = C.contract(ups, sep_vars(fe.psep[c]), false, ops);
  }
  C.saving = false;
  P.n_ups = (int)ups.size();
  const std::map<std::vector<int>, int> relayouts_after_ups = C.relayouts;
  // ---- forward program: the messages, then alpha_t (or, without an interface, the slice's mass) ----
  P.fwd = ups;
  {
    std::vector<FacOpRef> ops = L[fe.root];
    if (hm.nif > 0 && fe.root == hm.in_clique) ops.push_back(FacOpRef{t_alpha_in, false});
    for (int d : fe.children[fe.root]) ops.push_back(FacOpRef{up[d], false});
    const int t_alpha_new = C.contract(P.fwd, hm.nif > 0 ? hm.outg : std::vector<int>{}, true, ops);
    P.o_alpha_new = C.T[t_alpha_new].off;
    FacInstr s;
    s.kind = FI_SETTLE_FWD;
    P.fwd.push_back(s);
  }
  C.relayouts = relayouts_after_ups;
  // ---- backward program: targets, their host cliques ----
  struct Target { std::vector<int> vars; int kind, var; long long dst; bool only_t0; int clique; };
  std::vector<Target> targets;
  if (counts)
    for (int v = 0; v < hm.nv; v++) {
      Target t;
      t.vars = {v};
      for (int j = hm.poff[v]; j < hm.poff[v + 1]; j++) t.vars.push_back(hm.parents[j]);
      t.kind = FI_COUNT; t.var = v; t.dst = hm.coff[v];
      t.only_t0 = (hm.flags[v] & NIPGPU_IF_OLD_OUTGOING) != 0;
      t.clique = -1;
      targets.push_back(t);
    }
  if (queries) {
    long long off = 0;
    for (int q = 0; q < a.n_query; q++) {
      Target t;
      t.vars = {a.query[q]};
      t.kind = FI_QUERY; t.var = q; t.dst = off; t.only_t0 = false; t.clique = -1;
      off += hm.card[a.query[q]];
      targets.push_back(t);
    }
  }
  // Place every target (largest first) where it costs least.  A belief costs one pass over the
  // whole clique whatever it keeps (terms = clique entries), so targets gather in cliques that
  // already need a belief; a leaf clique gets its message from its parent's belief, so placing
  // a target there also makes the parent keep the sepset's variables.
  std::vector<std::vector<int>> U(nc);
  auto grown = [&](int c, const std::vector<int>& vars) {
    long long n = prod_card(hm, U[c]);
    for (int v : vars)
      if (!has_var(U[c], v)) n *= hm.card[v];
    return (double)n;
  };
  auto is_leaf_kid = [&](int c) {
    return c != fe.root && fe.children[c].empty() && !(hm.nif > 0 && c == hm.in_clique);
  };
  auto place_cost = [&](int c, const std::vector<int>& vars) {
    double cost = (U[c].empty() ? (double)hm.csize[c] : 0.0) + grown(c, vars);
    if (is_leaf_kid(c)) {
      const int p = fe.parent[c];
      cost += (U[p].empty() ? (double)hm.csize[p] : 0.0) + grown(p, sep_vars(fe.psep[c]));
    }
    return cost;
  };
  std::vector<int> order(targets.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    return prod_card(hm, targets[x].vars) > prod_card(hm, targets[y].vars);
  });
  for (int ti : order) {
    Target& t = targets[ti];
    double best = -1;
    for (int c = 0; c < nc; c++) {
      bool holds = true;
      for (int v : t.vars) holds = holds && hm.var_pos(c, v) >= 0;
      if (!holds) continue;
      const double g = place_cost(c, t.vars);
      const bool better = best < 0 || g < best || (g == best && c == hm.family[t.vars[0]]);
      if (better) { best = g; t.clique = c; }
    }
    if (t.clique < 0) { set_error("factor engine: a target is held by no clique"); return NIPGPU_EINVAL; }
    for (int v : t.vars)
      if (!has_var(U[t.clique], v)) U[t.clique].push_back(v);
    if (is_leaf_kid(t.clique))
      for (int v : sep_vars(fe.psep[t.clique]))
        if (!has_var(U[fe.parent[t.clique]], v)) U[fe.parent[t.clique]].push_back(v);
  }
  // which subtrees need a downward message
  std::vector<char> needed(nc, 0);
  for (int i = nc - 1; i >= 0; i--) {
    const int c = fe.preorder[i];
    bool n = !U[c].empty() || (hm.nif > 0 && c == hm.in_clique);
    for (int d : fe.children[c]) n = n || needed[d];
    needed[c] = n;
  }
  P.bwd = ups;
  for (int c : fe.preorder) {
    if (!needed[c]) continue;
    std::vector<FacOpRef> all = L[c];   // every operand of the clique's belief
    if (hm.nif > 0 && c == hm.in_clique) all.push_back(FacOpRef{t_alpha_in, false});
    if (hm.nif > 0 && c == fe.root) all.push_back(FacOpRef{t_beta, false});
    if (c != fe.root) all.push_back(FacOpRef{down[c], false});
    for (int d : fe.children[c]) all.push_back(FacOpRef{up[d], false});
    // leaf children that need a message get it from the belief (Hugin division by what they sent)
    std::vector<int> leaf_kids, inner_kids;
    for (int d : fe.children[c]) {
      if (!needed[d]) continue;
      const bool leaf = fe.children[d].empty() && !(hm.nif > 0 && d == hm.in_clique);
      (leaf ? leaf_kids : inner_kids).push_back(d);
    }
    for (int d : leaf_kids)
      for (int v : sep_vars(fe.psep[d]))
        if (!has_var(U[c], v)) U[c].push_back(v);
    int belief = -1;
    if (!U[c].empty()) {
      std::vector<int> uv;   // clique order
      for (int k = 0; k < hm.clique_dim(c); k++)
        if (has_var(U[c], hm.clique_vars(c)[k])) uv.push_back(hm.clique_vars(c)[k]);
      belief = C.contract(P.bwd, uv, false, all);
      // every marginal comes from the smallest tensor already computed here that holds its
      // variables (largest targets first: a 16-entry marginal is summed out of a 4096-entry
      // family table, not out of the belief again)
      std::vector<int> made{belief};
      auto smallest_source = [&](const std::vector<int>& vars) {
        int best = belief;
        for (int t : made) {
          bool holds = true;
          for (int v : vars) holds = holds && has_var(C.T[t].vars, v);
          if (holds && C.T[t].size < C.T[best].size) best = t;
        }
        return best;
      };
      std::vector<int> mine;
      for (size_t ti = 0; ti < targets.size(); ti++)
        if (targets[ti].clique == c) mine.push_back((int)ti);
      std::stable_sort(mine.begin(), mine.end(), [&](int x, int y) {
        return prod_card(hm, targets[x].vars) > prod_card(hm, targets[y].vars);
      });
      for (int ti : mine) {
        const Target& t = targets[ti];
        int src = smallest_source(t.vars);
        if (t.vars != C.T[src].vars) {
          src = C.contract(P.bwd, t.vars, true, {FacOpRef{src, false}});
          made.push_back(src);
        }
        FacInstr ins;
        ins.kind = t.kind;
        ins.src_off = C.T[src].off;
        ins.n = (int)C.T[src].size;
        ins.var = t.var;
        ins.dst_off = t.dst;
        ins.only_t0 = t.only_t0;
        P.bwd.push_back(ins);
      }
      for (int d : leaf_kids)
        down[d] = C.contract(P.bwd, sep_vars(fe.psep[d]), false,
                             {FacOpRef{smallest_source(sep_vars(fe.psep[d])), false}, FacOpRef{up[d], true}});
    }
    for (int d : inner_kids) {
      std::vector<FacOpRef> ops;
      for (const FacOpRef& o : all)
        if (o.tensor != up[d]) ops.push_back(o);
      down[d] = C.contract(P.bwd, sep_vars(fe.psep[d]), false, ops);
    }
    if (hm.nif > 0 && c == hm.in_clique) {   // message to slice t-1: everything but alpha_{t-1}
      std::vector<FacOpRef> ops;
      for (const FacOpRef& o : all)
        if (o.tensor != t_alpha_in) ops.push_back(o);
      const int t_bprev = C.contract(P.bwd, hm.prev, true, ops, true);
      P.o_bprev = C.T[t_bprev].off;
      FacInstr nb;
      nb.kind = FI_BETA_NORM;
      nb.needs_history = true;
      P.bwd.push_back(nb);
    }
  }
  if (C.failed) { set_error("factor engine: a contraction exceeds the engine's limits"); return NIPGPU_EUNSUPPORTED; }
  // partial sums of split contractions live behind the tensors
  P.o_partial = C.slot_top;
  P.partial_doubles = (C.partial_max + 15) / 16 * 16;
  P.slot_doubles = P.o_partial + P.partial_doubles + 16;
  P.saved_doubles = (C.saved_top + 15) / 16 * 16;
  for (std::vector<FacInstr>* v : {&P.fwd, &P.bwd})
    for (FacInstr& i : *v) {
      if (i.kind != FI_CONTRACT) continue;
      (v == &P.fwd ? P.flops_fwd : P.flops_bwd) += i.flops;
    }
  P.n_obs = a.n_obs;
  if (int e = dev_up(&P.d_pool, C.pool, st)) return e;
  if (int e = dev_up(&P.d_marked, marked, st)) return e;
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  return NIPGPU_OK;
}

template <int NSH, int NV, int TJ>
void launch_contract_n(const FacStepDev& s, const SlotCtx& X, int alive, cudaStream_t st) {
  const dim3 grid(s.cpc > 1 ? 1 : (s.n_thr + 127) / 128, (s.n_chunks + s.cpc - 1) / s.cpc, alive);
  const size_t smem = (size_t)std::max(NSH + NV, 1) * s.cpc * s.Rc * sizeof(int);
  bool ev = false;
  for (int k = 0; k < s.nSh + s.nVar; k++) ev = ev || s.opR[k].kind == FT_EVID;
  if (ev) k_fac_contract<NSH, NV, TJ, true><<<grid, 128, smem, st>>>(s, X);
  else k_fac_contract<NSH, NV, TJ, false><<<grid, 128, smem, st>>>(s, X);
}

template <int NSH, int TJ>
void launch_contract_v(const FacStepDev& s, const SlotCtx& X, int alive, cudaStream_t st) {
  if constexpr (TJ == 1) {
    launch_contract_n<NSH, 0, 1>(s, X, alive, st);
  } else {
    switch (s.nVar) {
      case 1: launch_contract_n<NSH, 1, TJ>(s, X, alive, st); break;
      case 2: launch_contract_n<NSH, 2, TJ>(s, X, alive, st); break;
      case 3: launch_contract_n<NSH, 3, TJ>(s, X, alive, st); break;
      default: launch_contract_n<NSH, 4, TJ>(s, X, alive, st); break;
    }
  }
}

template <int NSH, int NV, int NV2>
void launch_contract_2d(const FacStepDev& s, const SlotCtx& X, int alive, cudaStream_t st) {
  const dim3 grid(s.cpc > 1 ? 1 : (s.n_thr + 127) / 128, (s.n_chunks + s.cpc - 1) / s.cpc, alive);
  const size_t smem = (size_t)(NSH + NV + NV2) * s.cpc * s.Rc * sizeof(int);
  k_fac_contract2<NSH, NV, NV2, 4, 4><<<grid, 128, smem, st>>>(s, X);
}

void launch_contract(FacStepDev s, const SlotCtx& X, int alive, cudaStream_t st) {
  if (s.TK > 1) {   // 2-D register tile: 1..2 operands per tile variable, 0..2 shared (the planner checked)
    const int key = s.nSh * 100 + s.nVar * 10 + s.nVar2;
    switch (key) {
      case 11: launch_contract_2d<0, 1, 1>(s, X, alive, st); break;
      case 12: launch_contract_2d<0, 1, 2>(s, X, alive, st); break;
      case 21: launch_contract_2d<0, 2, 1>(s, X, alive, st); break;
      case 22: launch_contract_2d<0, 2, 2>(s, X, alive, st); break;
      case 111: launch_contract_2d<1, 1, 1>(s, X, alive, st); break;
      case 112: launch_contract_2d<1, 1, 2>(s, X, alive, st); break;
      case 121: launch_contract_2d<1, 2, 1>(s, X, alive, st); break;
      case 122: launch_contract_2d<1, 2, 2>(s, X, alive, st); break;
      case 211: launch_contract_2d<2, 1, 1>(s, X, alive, st); break;
      case 212: launch_contract_2d<2, 1, 2>(s, X, alive, st); break;
      case 221: launch_contract_2d<2, 2, 1>(s, X, alive, st); break;
      default: launch_contract_2d<2, 2, 2>(s, X, alive, st); break;
    }
    return;
  }
  if (s.TJ == 1 || s.nVar == 0) {
    // no tile (or nothing varies along it: every result of the tile equals the first one, still correct
    // to compute one by one): all in-loop operands are "shared"
    if (s.TJ != 1) {   // nVar == 0 with a tile: run it as TJ results per thread of the same sum
      switch (s.nSh) {
        case 0: launch_contract_n<0, 0, 4>(s, X, alive, st); break;
        case 1: launch_contract_n<1, 0, 4>(s, X, alive, st); break;
        case 2: launch_contract_n<2, 0, 4>(s, X, alive, st); break;
        case 3: launch_contract_n<3, 0, 4>(s, X, alive, st); break;
        default: launch_contract_n<4, 0, 4>(s, X, alive, st); break;
      }
      return;
    }
    switch (s.nSh) {
      case 0: launch_contract_n<0, 0, 1>(s, X, alive, st); break;
      case 1: launch_contract_n<1, 0, 1>(s, X, alive, st); break;
      case 2: launch_contract_n<2, 0, 1>(s, X, alive, st); break;
      case 3: launch_contract_n<3, 0, 1>(s, X, alive, st); break;
      case 4: launch_contract_n<4, 0, 1>(s, X, alive, st); break;
      case 5: launch_contract_n<5, 0, 1>(s, X, alive, st); break;
      case 6: launch_contract_n<6, 0, 1>(s, X, alive, st); break;
      case 7: launch_contract_n<7, 0, 1>(s, X, alive, st); break;
      default: launch_contract_n<8, 0, 1>(s, X, alive, st); break;
    }
    return;
  }
  switch (s.nSh) {
    case 0: launch_contract_v<0, 4>(s, X, alive, st); break;
    case 1: launch_contract_v<1, 4>(s, X, alive, st); break;
    case 2: launch_contract_v<2, 4>(s, X, alive, st); break;
    case 3: launch_contract_v<3, 4>(s, X, alive, st); break;
    default: launch_contract_v<4, 4>(s, X, alive, st); break;
  }
}

struct RunCtx {
  const FacProgram* P;
  const FacRunArgs* a;
  FacEngine* fe;
  const int* d_len_sorted;
  int S, nif, t_rows;
  bool no_beta_update;   // filtering: beta stays 1, no message to slice t-1
};

int run_instr(const RunCtx& R, const FacInstr& ins, const SlotCtx& X, int alive, cudaStream_t st) {
  const FacProgram& P = *R.P;
  const FacRunArgs& a = *R.a;
  switch (ins.kind) {
    case FI_CONTRACT: {
      if (ins.needs_history && (X.t == 0 || R.no_beta_update)) return NIPGPU_OK;
      FacStepDev s = ins.step;
      const long long final_off = s.out.off + (s.out.kind == FT_SAVED ? X.saved_off : 0);
      if (s.n_chunks > 1) { s.out.off = P.o_partial; s.out.kind = FT_SLOT; }
      launch_contract(s, X, alive, st);
      NIPGPU_LAUNCHED();
      if (s.n_chunks > 1) {
        const int per = s.n_chunks >= 32 ? 8 : 256;   // outputs per 256-thread CTA
        k_fac_reduce<<<dim3((s.n_out + per - 1) / per, alive), 256, 0, st>>>(X, P.o_partial, final_off, s.n_out,
                                                                             s.n_chunks);
        NIPGPU_LAUNCHED();
      }
      return NIPGPU_OK;
    }
    case FI_SETTLE_FWD:
      k_fac_settle_fwd<<<alive, 1024, 0, st>>>(X, R.d_len_sorted, P.d_marked, R.S, P.o_alpha_in, P.o_alpha_new,
                                               a.d_R1, a.d_m10, R.fe->d_alpha_wave, R.t_rows, a.want_ll, R.nif, R.fe->d_ll_run,
                                               R.fe->d_bad_run, a.d_ll, a.d_status);
      NIPGPU_LAUNCHED();
      return NIPGPU_OK;
    case FI_BETA_NORM:
      if (X.t == 0 || R.no_beta_update) return NIPGPU_OK;
      k_fac_beta_norm<<<alive, 1024, 0, st>>>(X, R.S, P.o_bprev, P.o_beta);
      NIPGPU_LAUNCHED();
      return NIPGPU_OK;
    case FI_COUNT:
      if (!a.d_acc || (ins.only_t0 && X.t > 0)) return NIPGPU_OK;
      k_fac_count<<<alive, 1024, 0, st>>>(X, ins.src_off, ins.n, a.d_acc, a.acc_stride, ins.dst_off);
      NIPGPU_LAUNCHED();
      return NIPGPU_OK;
    case FI_QUERY:
      if (!a.d_post) return NIPGPU_OK;
      k_fac_query<<<alive, 256, 0, st>>>(X, ins.src_off, ins.n, a.d_post, a.post_row, ins.dst_off);
      NIPGPU_LAUNCHED();
      return NIPGPU_OK;
  }
  return NIPGPU_OK;
}

}  // namespace

#ifndef NIPGPU_FACTOR_MAX_SLOTS
#define NIPGPU_FACTOR_MAX_SLOTS 64
#endif
int fac_slots(const HostModel& hm, const FacEngine& fe, int n_series) {
  (void)fe;
  static const int forced = [] {
    const char* p = getenv("NIPGPU_FACTOR_SLOTS");
    return p ? atoi(p) : 0;
  }();
  // sequences in flight: NIPGPU_FACTOR_MAX_SLOTS for models whose messages are megabytes (one
  // contraction of one sequence already fills the machine), more for small models, where the
  // launches of a slice have to be shared by many sequences: about 2 GB of work areas
  const double per_slot = 8.0 * (4.0 * hm.msg_total + 6.0 * hm.S + 1024.0);
  long long w = (long long)(2147483648.0 / per_slot);
  w = std::max<long long>(NIPGPU_FACTOR_MAX_SLOTS, std::min<long long>(w, 4096));
  if (forced > 0) w = forced;
  return (int)std::max<long long>(1, std::min<long long>(w, n_series));
}

int fac_run(const HostModel& hm, FacEngine& fe, const FacRunArgs& a, cudaStream_t st) {
  if (!fe.ok) { set_error("factor engine: " + fe.why); return NIPGPU_EUNSUPPORTED; }
  if (a.n_series == 0) return NIPGPU_OK;
  const bool counts = a.d_acc != nullptr;
  const bool queries = a.d_post && a.n_query > 0;
  const std::vector<int> key = plan_key(a, hm, counts);
  auto it = fe.programs.find(key);
  if (it == fe.programs.end()) {
    FacProgram P;
    if (int e = compile(hm, fe, a, counts, P, st)) { cudaFree(P.d_pool); cudaFree(P.d_marked); return e; }
    it = fe.programs.emplace(key, P).first;
  }
  const FacProgram& P = it->second;
  // ---- sequences sorted by length, longest first; W of them in flight ----
  std::vector<int> order(a.n_series);
  std::iota(order.begin(), order.end(), 0);
  const std::vector<int>& len = *a.len;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len[x] > len[y]; });
  std::vector<int> len_sorted(a.n_series);
  for (int i = 0; i < a.n_series; i++) len_sorted[i] = len[order[i]];
  int W = fac_slots(hm, fe, a.n_series);
  if (counts) W = std::min(W, a.acc_slots);
  // the upward messages of every slice are kept for the backward pass when that fits
  const bool backward_pass = !a.forward_only && (queries || counts) && hm.nif > 0;
  const size_t budget = (size_t)24 << 30;
  const int t_longest = std::max(1, len_sorted[0]);
  bool keep_ups = backward_pass;
  {
    static const bool off = [] { const char* p = getenv("NIPGPU_FACTOR_KEEP_UPS"); return p && p[0] == '0'; }();
    const size_t full = (size_t)(P.slot_doubles + P.saved_doubles * t_longest) * sizeof(double);
    if (off || (size_t)std::min(W, 16) * full > budget) keep_ups = false;
  }
  const long long slot_doubles = P.slot_doubles + P.saved_doubles * (keep_ups ? t_longest : 1);
  const size_t slot_bytes = (size_t)slot_doubles * sizeof(double);
  const size_t row_bytes = (size_t)t_longest * std::max(hm.S, 1) * sizeof(double);   // its forward rows
  while (W > 1 && (size_t)W * (slot_bytes + row_bytes) > budget) W--;
  if (fe.slots_cap < (size_t)W * slot_doubles) {
    cudaFree(fe.d_slots);
    fe.d_slots = nullptr;
    fe.slots_cap = 0;
    NIPGPU_CUDA(cudaMalloc((void**)&fe.d_slots, (size_t)W * slot_bytes));
    fe.slots_cap = (size_t)W * slot_doubles;
  }
  {   // forward rows of the sequences in flight
    const size_t need = (size_t)W * t_longest * std::max(hm.S, 1);
    if (fe.alpha_cap < need) {
      cudaFree(fe.d_alpha_wave);
      fe.d_alpha_wave = nullptr;
      fe.alpha_cap = 0;
      NIPGPU_CUDA(cudaMalloc((void**)&fe.d_alpha_wave, need * sizeof(double)));
      fe.alpha_cap = need;
    }
  }
  if (fe.max_slots < W) {
    cudaFree(fe.d_ll_run); cudaFree(fe.d_bad_run);
    fe.d_ll_run = nullptr; fe.d_bad_run = nullptr;
    NIPGPU_CUDA(cudaMalloc((void**)&fe.d_ll_run, W * sizeof(double)));
    NIPGPU_CUDA(cudaMalloc((void**)&fe.d_bad_run, W * sizeof(int)));
    fe.max_slots = W;
  }
  int *d_order = nullptr, *d_len_sorted = nullptr;
  if (int e = dev_up(&d_order, order, st)) return e;
  if (int e = dev_up(&d_len_sorted, len_sorted, st)) { cudaFree(d_order); return e; }
  auto done = [&](int rc) {
    cudaStreamSynchronize(st);
    cudaFree(d_order);
    cudaFree(d_len_sorted);
    return rc;
  };
  RunCtx R{&P, &a, &fe, d_len_sorted, hm.S, hm.nif, t_longest, false};
  // NIPGPU_FACTOR_TRACE=1: device time of every instruction of the first wave's slices (stderr)
  static const bool trace = [] { const char* p = getenv("NIPGPU_FACTOR_TRACE"); return p && p[0] == '1'; }();
  std::vector<double> tr_f(P.fwd.size(), 0.0), tr_b(P.bwd.size(), 0.0);
  cudaEvent_t te0 = nullptr, te1 = nullptr;
  if (trace) { cudaEventCreate(&te0); cudaEventCreate(&te1); }
  auto timed = [&](const FacInstr& ins, const SlotCtx& Xc, int alive, double* slot) -> int {
    if (!trace) return run_instr(R, ins, Xc, alive, st);
    cudaEventRecord(te0, st);
    const int e = run_instr(R, ins, Xc, alive, st);
    cudaEventRecord(te1, st);
    cudaEventSynchronize(te1);
    float ms = 0;
    cudaEventElapsedTime(&ms, te0, te1);
    *slot += ms;
    return e;
  };
  SlotCtx X;
  X.slots = fe.d_slots; X.slot_stride = slot_doubles; X.fac = fe.d_fac; X.pool = P.d_pool;
  X.saved_off = P.slot_doubles;
  X.obs = a.d_obs; X.n_obs = a.n_obs; X.row_off = a.d_row_off; X.order = d_order;
  const long long a0_off = fe.a0_tensor >= 0 ? fe.tensors[fe.a0_tensor].off : 0;
  const bool filtered = a.forward_only && queries;
  const bool backward = !a.forward_only && (queries || counts);
  const int S = hm.S;
  const int prep_blocks = std::max(1, std::min(64, (S + 255) / 256));
  auto alive_at = [&](int wave0, int nslots, int t) {
    int n = 0;
    while (n < nslots && len_sorted[wave0 + n] > t) n++;
    return n;
  };
  for (int wave0 = 0; wave0 < a.n_series; wave0 += W) {
    const int nslots = std::min(W, a.n_series - wave0);
    const int Tw = len_sorted[wave0];
    X.wave0 = wave0;
    for (int t = 0; t < Tw; t++) {
      const int alive = alive_at(wave0, nslots, t);
      if (alive == 0) break;
      X.t = t;
      X.saved_off = P.slot_doubles + (keep_ups ? (long long)t * P.saved_doubles : 0);
      k_fac_prepare<<<dim3(prep_blocks, alive), 256, 0, st>>>(X, d_len_sorted, fe.d_alpha_wave, t_longest, hm.nif > 0 ? S : 0,
                                                              P.o_alpha_in, P.o_beta, a0_off, filtered ? 2 : 0,
                                                              fe.d_ll_run, fe.d_bad_run, t == 0);
      NIPGPU_LAUNCHED();
      R.no_beta_update = false;
      for (size_t i = 0; i < P.fwd.size(); i++)
        if (int e = timed(P.fwd[i], X, alive, &tr_f[i])) return done(e);
      if (filtered || (hm.nif == 0 && (queries || counts))) {
        // filtered marginals (or a model without an interface): beliefs of this slice alone
        R.no_beta_update = true;    // (the upward messages of this slice were just computed)
        for (size_t i = P.n_ups; i < P.bwd.size(); i++)
          if (int e = timed(P.bwd[i], X, alive, &tr_b[i])) return done(e);
      }
    }
    if (backward && hm.nif > 0) {
      R.no_beta_update = false;
      for (int t = Tw - 1; t >= 0; t--) {
        const int alive = alive_at(wave0, nslots, t);
        if (alive == 0) continue;
        X.t = t;
        X.saved_off = P.slot_doubles + (keep_ups ? (long long)t * P.saved_doubles : 0);
        k_fac_prepare<<<dim3(prep_blocks, alive), 256, 0, st>>>(X, d_len_sorted, fe.d_alpha_wave, t_longest, S, P.o_alpha_in, P.o_beta,
                                                                a0_off, 1, fe.d_ll_run, fe.d_bad_run, 0);
        NIPGPU_LAUNCHED();
        for (size_t i = keep_ups ? P.n_ups : 0; i < P.bwd.size(); i++)
          if (int e = timed(P.bwd[i], X, alive, &tr_b[i])) return done(e);
      }
    }
  }
  if (trace) {
    auto dump = [&](const char* name, const std::vector<FacInstr>& v, const std::vector<double>& ms) {
      for (size_t i = 0; i < v.size(); i++) {
        const FacStepDev& s = v[i].step;
        if (v[i].kind == FI_CONTRACT)
          fprintf(stderr, "[factor] %s %2zu contract n_out %8d R %6d chunks %4d TJ %dx%d nSh %d nVar %d+%d nO %d  %9.3f ms  %.2f GFLOP\n", name, i,
                  s.n_out, s.R, s.n_chunks, s.TJ, s.TK, s.nSh, s.nVar, s.nVar2, s.nO, ms[i], v[i].flops * 1e-9);
        if (v[i].kind == FI_CONTRACT && ms[i] > 1.0) fprintf(stderr, "          %s\n", v[i].note.c_str());
        else
          fprintf(stderr, "[factor] %s %2zu kind %d n %d  %9.3f ms\n", name, i, v[i].kind, v[i].n, ms[i]);
      }
    };
    dump("fwd", P.fwd, tr_f);
    dump("bwd", P.bwd, tr_b);
    cudaEventDestroy(te0);
    cudaEventDestroy(te1);
  }
  return done(NIPGPU_OK);
}

}  // namespace nipgpu
