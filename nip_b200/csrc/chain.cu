// chain.cu — engine 2, see chain.cuh.  FP64 tensor-core (DMMA) forward /
// backward recursions for interface-clique + leaf models.
//
// Fragment convention (mma.sync.aligned.m8n8k4 .f64, lane L, g = L/4, q = L%4):
//   A (8x4, row)  : lane holds A[g][q]
//   B (4x8, col)  : lane holds B[q][g]
//   C/D (8x8)     : lane holds C[g][2q], C[g][2q+1]
// A warp keeps a [8 sequences] x [SP states] vector as NT = SP/8 accumulator
// tiles, i.e. lane (g,q) owns states 8n+2q+e (n < NT, e < 2) of sequence g.
// The K loop of the next contraction enumerates the states in exactly that
// ownership order (k-step (n,e), k-slot q  <->  state 8n+2q+e), so the previous
// result is already the A operand: no shuffle, no shared-memory round trip.
// The transition table is pre-arranged once per parameter change in "fragment
// order" so that every B operand is one conflict-free 8-byte load per lane.
#include "chain.cuh"
#include "sweep.cuh"

#include <algorithm>
#include <cfloat>
#include <type_traits>

namespace nipgpu {

namespace {

// position of A(k-state, n-state) inside a fragment-ordered SPxSP table
// For NT >= 2 the fragments of n-tiles (2m, 2m+1) of one lane are adjacent (one LDS.128).
__host__ __device__ inline int frag_index(int kstate, int nstate, int NT) {
  const int j = kstate >> 3, q = (kstate & 7) >> 1, e = kstate & 1;
  const int nt = nstate >> 3, g = nstate & 7;
  const int ks = j * 2 + e, lane = (g << 2) + q;
  if (NT >= 2) return ((((ks * (NT / 2) + (nt >> 1)) << 5) + lane) << 1) + (nt & 1);
  return ((ks * NT + nt) << 5) + lane;
}

struct ChainDev {
  int S, SP, c_miss;   // c_miss: combined evidence index meaning "no evidence in this slice"
  int AS;              // doubles per stored forward row: SP, or S rounded up to even when one 8-state
                       // tile holds the interface (a 4-state model moves 32 B per row, not 64)
  double m1_0;         // mass of the evidence-free first slice
  const double *Bf1, *Bb1, *Bb0, *phi0, *lam0, *R1, *colsum, *lam_comb;
};

struct ChainBatchDev {
  int* fexp;               // [rows] exponent carried by the stored forward row (pair kernels)
  double* zc;              // [n_series] sum of the last forward row
  int* zf;                 // [n_series] its exponent
  double* rn_out;          // [rows] E-step: 1 / (forward row . beta row), or nullptr
  int n_series;
  const int* order;        // sorted position -> series
  const int* len_sorted;
  const int* cfg;          // [rows] combined evidence index, API row order
  const long long* row_off;
};

struct ChainComb {
  int mult[8], n_cfg[8];
  long long lam_off[8];
};

// ---------------------------------------------------------------- refresh ---
__global__ void k_chain_mats(const double* base0, const double* base1, const int* ent_of, int S,
                             int NT, double* Bf1, double* Bb1, double* Bb0) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= S * S) return;
  const int im = x / S, ip = x - im * S, ent = ent_of[x];
  const double a1 = base1[ent], a0 = base0[ent];
  Bf1[frag_index(im, ip, NT)] = a1;  // forward:  k = previous state, n = current state
  Bb1[frag_index(ip, im, NT)] = a1;  // backward: k = current state,  n = previous state
  Bb0[frag_index(ip, im, NT)] = a0;
}

__global__ void k_chain_phi0(const double* base0, const int* ent_of, int S, double* phi0) {
  // a warp per current state: the lanes share the sum over the previous states
  const int ip = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ip >= S) return;
  double s = 0;
  for (int im = lane; im < S; im += 32) s += base0[ent_of[im * S + ip]];
  s = warp_sum(s);
  if (lane == 0) phi0[ip] = s;
}

// meta: [n_free, card[0..n_free), stride[0..n_free)] — Lambda[cfg][ip] = sum over the leaf's
// free-variable combinations compatible with cfg of leaf_table[base[s(ip)] + off[r]]
__global__ void k_chain_lambda(const double* leaf_tab, const int* base, const int* off, int R,
                               const int* ip_to_s, const int* meta, int n_cfg, int S, int SP,
                               double* lam) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_cfg * S) return;
  const int cfg = x / S, ip = x - cfg * S;
  const int nf = meta[0];
  const int* card = meta + 1;
  const int* stride = meta + 1 + nf;
  const int b = base[ip_to_s[ip]];
  double s = 0;
  for (int r = 0; r < R; r++) {
    int rem = r, ok = 1;
    for (int k = 0; k < nf; k++) {
      const int digit = rem % card[k];
      rem /= card[k];
      const int code = (cfg / stride[k]) % (card[k] + 1);
      if (code != card[k] && code != digit) ok = 0;
    }
    if (ok) s += leaf_tab[b + off[r]];
  }
  lam[(long long)cfg * SP + ip] = s;
}

// out[ip] = prod over the listed leaves of Lambda_l[miss_l][ip]
__global__ void k_chain_lam_prod(const double* lam, const long long* row_off, int n, int S, int SP,
                                 double* out) {
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= SP) return;
  double p = ip < S ? 1.0 : 0.0;
  for (int l = 0; l < n && ip < S; l++) p *= lam[row_off[l] + ip];
  out[ip] = p;
}

// Combined evidence table of one call: Lc[c][ip] = lam_static[ip] * prod_a Lambda_a[digit_a(c)][ip]
// with c = sum_a cfg_a * mult_a.  One gather per (sequence, slice) in the hot kernels.
__global__ void k_chain_combine(const double* lam, const double* lam_static, int n_active,
                                ChainComb K, int n_comb, int S, int SP, double* Lc) {
  const long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (x >= (long long)n_comb * SP) return;
  const int c = (int)(x / SP), ip = (int)(x - (long long)c * SP);
  double p = ip < S ? lam_static[ip] : 0.0;
  for (int a = 0; a < n_active && ip < S; a++) {
    const int digit = (c / K.mult[a]) % K.n_cfg[a];
    p *= lam[K.lam_off[a] + (long long)digit * SP + ip];
  }
  Lc[x] = p;
}

// combined evidence configuration of every data row (API row order)
__global__ void k_chain_cfg(const int* obs, long long rows, int n_obs, const int* col_slot,
                            const int* col_stride, const int* col_card, const int* col_mult,
                            int c_miss, int* cfg) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int c = c_miss;
  for (int k = 0; k < n_obs; k++) {
    const int o = obs[r * n_obs + k];
    if (col_slot[k] >= 0 && o >= 0) c += (o - col_card[k]) * col_stride[k] * col_mult[k];
  }
  cfg[r] = c;
}

// running log-likelihood sum_t log(m2_t) - log(m1_t) kept as a product with a
// separate binary exponent: two multiplies per slice instead of two logarithms.
struct LogAcc {
  double p1 = 1.0, p2 = 1.0;  // mantissa products of the m1 / m2 factors
  int e1 = 0, e2 = 0;         // their binary exponents
  int zero = 0, bad = 0;
  __device__ __forceinline__ static void renorm(double& p, int& e) {
    const int hi = __double2hiint(p);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    e += ex;
    p = __hiloint2double(hi - (ex << 20), __double2loint(p));
  }
  // semantics of src/nip.c:1458-1474 and of the BAD_LUCK test :1827-1831; m1 and m2 may
  // carry a common positive factor (only their ratio enters the sum)
  __device__ __forceinline__ void add(double m1, double m2, bool on) {
    const bool both = on && m1 > 0 && m2 > 0;
    p1 *= both ? m1 : 1.0;
    p2 *= both ? m2 : 1.0;
    renorm(p1, e1);
    renorm(p2, e2);
    if (on && m2 == 0) zero = 1;
    const bool pos = e2 > e1 || (e2 == e1 && p2 > p1);  // running log-likelihood > 0
    if (on && (m1 <= 0 || m2 <= 0 || (pos && !zero))) bad = 1;
  }
  __device__ __forceinline__ double value() const {
    if (zero) return -DBL_MAX;
    return (log(p2) - log(p1)) + (double)(e2 - e1) * 0.693147180559945309417232121458;
  }
};

// 16-byte read-only global load that ptxas may not sink towards its use: it has to be
// issued where it is written (before / early in a sweep) to have its latency hidden.
__device__ __forceinline__ double2 ldg_pinned(const double2* p, bool on) {
  double2 v = make_double2(0.0, 0.0);
  if (on) asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ double quad_sum_full(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// 1/x for the positive, normal-range scale factors of the recursions (0 -> 1: a zero vector
// stays zero, nip_normalise_array).  Hardware seed + two Newton steps: branch-free, so it
// does not split the sweep's basic block the way the IEEE division's slow path would;
// relative error ~2e-16, far inside the 1e-9 gate.
__device__ __forceinline__ double safe_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  return x != 0 ? y : 1.0;
}

// ---------------------------------------------------------------- forward ---
// alpha_t = normalise((alpha_{t-1} . A) * lambda_t) and the likelihood terms.
//
// The recurrence is sequential in t and a warp is alone on its scheduler, so
// everything between two MMA sweeps is exposed latency.  Only ONE multiply per
// element is left there: the vector carried from slice to slice is
//     own_t = (own_{t-1} . A) * lambda_t * g_t
// where g_t is a scale factor computed DURING sweep t from own_{t-1} alone
// (g_t = 1 / (c_{t-1} m2_{t-1}), c = sum of own), so own stays O(1) without a
// reduction on the critical path.  alpha_t is own_t up to the scale c_t; the
// exact masses follow one slice later, inside the next sweep's issue gaps:
//     m2_t = c_t / (c_{t-1} g_t),      m1_t = (own_{t-1} . R1) / c_{t-1}.
// The alpha rows stored for the backward pass keep the scale c_t (the backward
// pass normalises every posterior by its own sum, as the reference does).
template <int NT, bool FILT, bool WLL>
__global__ void __launch_bounds__(128, 1) k_chain_forward(ChainDev C, ChainBatchDev B,
                                                          double* __restrict__ alpha,
                                                          double* __restrict__ post, int post_stride,
                                                          int post_off, double* ll_out,
                                                          int* status_out) {
  constexpr int SP = 8 * NT;
  extern __shared__ double sB[];
  double* s_r1 = sB + SP * SP;
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) sB[i] = C.Bf1[i];
  for (int i = threadIdx.x; i < SP; i += blockDim.x) s_r1[i] = C.R1[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int bp = warp * 8 + g;
  const bool valid = bp < B.n_series;
  const int T = valid ? B.len_sorted[bp] : 0;
  const int Tw = __shfl_sync(0xffffffffu, T, 0);  // sorted by length: row 0 is the longest
  const int orig = valid ? B.order[bp] : 0;
  const long long row0 = valid ? B.row_off[orig] : 0;
  const int* cfg = B.cfg + row0;
  const double* frag = sB + (NT >= 2 ? 2 * lane : lane);
  const double2* r1v = reinterpret_cast<const double2*>(s_r1);

  double own[NT][2], acc[NT][2], lam[NT][2];
  LogAcc L;
  auto load_lam = [&](int c, bool on) {
    const double2* p = reinterpret_cast<const double2*>(C.lam_comb + (long long)c * SP);
#pragma unroll
    for (int n = 0; n < NT; n++) {
      const double2 v = ldg_pinned(p + 4 * n + q, on);
      lam[n][0] = v.x;
      lam[n][1] = v.y;
    }
  };
  // what is still owed for the slice whose vector sits in `own`
  double K = 0;          // m2 of that slice = (sum of own) * K
  double Pp = 0;         // its m1 numerator (already times c of the slice before)
  double cp = 1.0;       // c of the slice before it
  bool noev_p = false, on_p = false;

  // ---- slice 0: own_0 = phi0 * lambda_0 / S0 (c_0 = 1), m1_0 = mass of the evidence-free slice
  int c_cur = T > 0 ? cfg[0] : 0;
  int c_next = T > 1 ? cfg[1] : 0;
  load_lam(c_cur, T > 0);
  {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int n = 0; n < NT; n++) {
      acc[n][0] = C.phi0[8 * n + 2 * q] * lam[n][0];
      acc[n][1] = C.phi0[8 * n + 2 * q + 1] * lam[n][1];
      s0 += acc[n][0];
      s1 += acc[n][1];
    }
    const double S0 = quad_sum_full(s0 + s1);
    const double inv = safe_rcp(S0);
#pragma unroll
    for (int n = 0; n < NT; n++) { own[n][0] = acc[n][0] * inv; own[n][1] = acc[n][1] * inv; }
    K = S0; Pp = C.m1_0; cp = 1.0; noev_p = c_cur == C.c_miss; on_p = T > 0;
  }

  // Work of one slice that is NOT on the recurrence's critical path, cut into items that are
  // hung between the tensor instructions of the next sweep (or run back to back after the last
  // slice).  It settles slice s = t-1 whose vector is the sweep's A operand `own`.
  constexpr int E = 2 * NT;                  // elements of a vector held by one lane
  constexpr int I_SUM = 0;                   // E items : partial sums of own
  constexpr int I_RED = I_SUM + E;           // 3 items : combine + two quad shuffles -> c
  constexpr int I_LL = I_RED + 3;            // 1 item  : masses, likelihood bookkeeping
  constexpr int I_ST = I_LL + 1;             // NT items: alpha row (and filtered output)
  constexpr int I_DOT = I_ST + NT;           // NT items: own . R1 (m1 numerator of the next slice)
  constexpr int I_DRED = I_DOT + NT;         // 3 items
  constexpr int I_G = I_DRED + 3;            // 1 item  : scale of the sweep in progress
  constexpr int I_LAM = I_G + 1;             // NT items: fold it into the evidence row
  constexpr int W_ITEMS = I_LAM + NT;
  double ps[4], cs = 0, m2p = 0, cinv = 1.0, d0 = 0, d1 = 0, Pn = 0, gscale = 1.0;
  int s_slice = 0;
  bool fold = true;                          // false for the epilogue after the last slice
  auto item = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w >= I_SUM && w < I_RED) {
      constexpr int i = w - I_SUM;
      const double x = own[i >> 1][i & 1];
      if constexpr (i < 4) ps[i & 3] = x;
      else ps[i & 3] += x;
    } else if constexpr (w == I_RED) {
      if constexpr (E >= 4) cs = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      else cs = ps[0] + ps[1];
    } else if constexpr (w == I_RED + 1) {
      cs += __shfl_xor_sync(0xffffffffu, cs, 1);
    } else if constexpr (w == I_RED + 2) {
      cs += __shfl_xor_sync(0xffffffffu, cs, 2);
    } else if constexpr (w == I_LL) {
      m2p = cs * K;
      if (WLL) L.add(Pp, noev_p ? Pp : m2p * cp, on_p);
      if (FILT) cinv = safe_rcp(cs);
    } else if constexpr (w >= I_ST && w < I_DOT) {
      constexpr int n = w - I_ST;
      const int AS = NT == 1 ? C.AS : SP;
      if (on_p && (NT > 1 || 2 * q < AS))
        reinterpret_cast<double2*>(alpha + (row0 + s_slice) * AS)[4 * n + q] = make_double2(own[n][0], own[n][1]);
      if (FILT) {  // filtering: the forward marginal of I_s is alpha_s = own / c
        double* prow = post + (row0 + s_slice) * post_stride + post_off;
        const int col = 8 * n + 2 * q;
        if (on_p && col < C.S) prow[col] = own[n][0] * cinv;
        if (on_p && col + 1 < C.S) prow[col + 1] = own[n][1] * cinv;
      }
    } else if constexpr (w >= I_DOT && w < I_DRED) {
      if (WLL && fold) {
        constexpr int n = w - I_DOT;
        const double2 v = r1v[4 * n + q];
        if constexpr (n == 0) { d0 = own[n][0] * v.x; d1 = own[n][1] * v.y; }
        else { d0 += own[n][0] * v.x; d1 += own[n][1] * v.y; }
      }
    } else if constexpr (w == I_DRED) {
      if (WLL && fold) Pn = d0 + d1;
    } else if constexpr (w == I_DRED + 1) {
      if (WLL && fold) Pn += __shfl_xor_sync(0xffffffffu, Pn, 1);
    } else if constexpr (w == I_DRED + 2) {
      if (WLL && fold) Pn += __shfl_xor_sync(0xffffffffu, Pn, 2);
    } else if constexpr (w == I_G) {
      if (fold) {
        const double den = cs * m2p;
        gscale = safe_rcp(den);
        K = den != 0 ? m2p : 0.0;
      }
    } else if constexpr (w >= I_LAM && w < W_ITEMS) {
      if (fold) {
        constexpr int n = w - I_LAM;
        lam[n][0] *= gscale;
        lam[n][1] *= gscale;
      }
    }
  };

  // The evidence row of slice t is requested at the END of iteration t-1 (ptxas sinks a load
  // towards its first use but not across the loop's back edge), so it is in flight during
  // the whole sweep of slice t and is consumed only by the sweep's last side items.
  c_cur = c_next;
  load_lam(c_cur, 1 < T);
  if (2 < T) c_next = __ldg(cfg + 2);
  for (int t = 1; t < Tw; t++) {
    const bool on = t < T;
    s_slice = t - 1;
    mma_sweep<NT>(acc, own, frag, [&](auto sc) { run_items<2 * NT * NT, W_ITEMS, decltype(sc)::value>(item); });
    Pp = Pn; cp = cs; noev_p = c_cur == C.c_miss; on_p = on;
#pragma unroll
    for (int n = 0; n < NT; n++) { own[n][0] = acc[n][0] * lam[n][0]; own[n][1] = acc[n][1] * lam[n][1]; }
    c_cur = c_next;
    load_lam(c_cur, t + 1 < T);     // evidence row of slice t+1
    if (t + 2 < T) c_next = __ldg(cfg + t + 2);
  }
  fold = false;
  s_slice = Tw - 1;
  static_for<0, W_ITEMS>(item);
  if (valid && q == 0) {
    if (ll_out) ll_out[orig] = (WLL && T > 0) ? L.value() : 0.0;
    if (status_out) status_out[orig] = L.bad;
  }
}

// --------------------------------------------------------------- backward ---
// Scaled backward recursion.  With beta_t = gamma_{t+1} / alpha_t (the ratio the
// reference multiplies into out_clique, src/nip.c:1518-1529) one has
//     beta_{t-1}  proportional to  A . (lambda_t * beta_t)
//     P(I_t | all evidence) = normalise(alpha_t * beta_t)
// so the per-element division of the literal schedule disappears; where alpha_t
// is 0 the posterior is 0 whatever beta_t holds, which is the reference's
// 0/0 -> 0 rule (src/nippotential.c:486-491).
// Between two sweeps only two multiplies per element remain: the scale of
// beta_{t-1} (1 / sum of the sweep's result) is obtained during the sweep as
// 1 / (r . colsum(A)), and the posterior of slice t, the alpha / evidence
// prefetches and the stores all sit in the sweep's issue gaps.
// EM variant: stores, per slice, only the scaled beta_t it carries (rt[t]) and the scale h_t
// with beta_{t-1} = h_t A r_t (hvec[t]); it neither reads the forward store nor forms
// posteriors.  Everything the E-step needs follows from (own_t, beta_t, h_t, evidence index)
// row by row in k_chain_stats: posterior_t = own_t beta_t / N_t with N_t = own_t . beta_t, and
// the slice's joint mass Z_t = own_{t-1} . (A r_t) = N_{t-1} / h_t.  r0[series] =
// r_0 / (phi0 . r_0) is still written here for the first slice.
template <int NT, bool VEC, bool EM>
__global__ void __launch_bounds__(128, 1) k_chain_backward(ChainDev C, ChainBatchDev B,
                                                           const double* __restrict__ alpha,
                                                           double* __restrict__ post,
                                                           int post_stride, int post_off,
                                                           double* __restrict__ rt,
                                                           double* __restrict__ r0,
                                                           double* __restrict__ hvec) {
  constexpr int SP = 8 * NT;
  extern __shared__ double sB[];
  double* s_cs = sB + SP * SP;
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) sB[i] = C.Bb1[i];
  for (int i = threadIdx.x; i < SP; i += blockDim.x) s_cs[i] = C.colsum[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int bp = warp * 8 + g;
  const bool valid = bp < B.n_series;
  const int T = valid ? B.len_sorted[bp] : 0;
  const int Tw = __shfl_sync(0xffffffffu, T, 0);
  const int orig = valid ? B.order[bp] : 0;
  const long long row0 = valid ? B.row_off[orig] : 0;
  const int* cfg = B.cfg + row0;
  const double* frag = sB + (NT >= 2 ? 2 * lane : lane);
  const double2* csv = reinterpret_cast<const double2*>(s_cs);
  double beta[NT][2], r[NT][2], lam[NT][2], a[NT][2], an[NT][2], u[NT][2];
  const int AS = NT == 1 ? C.AS : SP;                 // stride of the stored forward rows
  const bool arow = NT > 1 || 2 * q < AS;             // this lane's pair of states exists in them
  auto load_row = [&](const double* base, bool on, double (&dst)[NT][2]) {
    const double2* p = reinterpret_cast<const double2*>(base);
#pragma unroll
    for (int n = 0; n < NT; n++) {
      const double2 v = ldg_pinned(p + 4 * n + q, on);
      dst[n][0] = v.x;
      dst[n][1] = v.y;
    }
  };
  // prologue: the longest rows start at slice Tw-1 with beta = 1, r = lambda
  const bool has_last = Tw >= 1 && Tw - 1 < T;
  load_row(C.lam_comb + (long long)(has_last ? cfg[Tw - 1] : 0) * SP, has_last, r);
  if (!EM) load_row(alpha + (row0 + Tw - 1) * AS, has_last && arow, a);
#pragma unroll
  for (int n = 0; n < NT; n++) beta[n][0] = beta[n][1] = 1.0;
  int c_pre = (Tw >= 2 && Tw - 2 < T) ? cfg[Tw - 2] : 0;

  // Work of slice t that is not on the recurrence's critical path, cut into items hung
  // between the tensor instructions of sweep t (see mma_sweep / run_items).
  constexpr int E = 2 * NT;
  constexpr int I_MUL = 0;                 // E items : a = alpha_t * beta_t, partial sums
  constexpr int I_RED = I_MUL + E;         // 3 items : combine + two quad shuffles
  constexpr int I_INV = I_RED + 3;         // 1 item  : reciprocal of the posterior's sum
  constexpr int I_ST = I_INV + 1;          // NT items: store the posterior of slice t
  constexpr int I_LD = I_ST + NT;          // NT items: alpha_{t-1} (prefetched into `an` before the sweep) -> `a`
  constexpr int I_DOT = I_LD + NT;         // NT items: r . colsum(A) = sum of the sweep's result
  constexpr int I_DRED = I_DOT + NT;       // 3 items
  constexpr int I_H = I_DRED + 3;          // 1 item  : scale of beta_{t-1}
  constexpr int I_LAM = I_H + 1;           // NT items: fold it into the evidence row of slice t-1
  constexpr int W_ITEMS = I_LAM + NT;
  double ps[4], psum = 0, pinv = 1.0, d0 = 0, d1 = 0, dsum = 0, h = 1.0, hs = 1.0;
  int t_cur = 0;
  bool on = false, pre = false, first_next = false;
  auto item = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w >= I_MUL && w < I_RED) {
      if constexpr (!EM) {
        constexpr int i = w - I_MUL;
        a[i >> 1][i & 1] *= beta[i >> 1][i & 1];
        if constexpr (i < 4) ps[i & 3] = a[i >> 1][i & 1];
        else ps[i & 3] += a[i >> 1][i & 1];
      }
    } else if constexpr (w == I_RED) {
      if constexpr (!EM) {
        if constexpr (E >= 4) psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
        else psum = ps[0] + ps[1];
      }
    } else if constexpr (w == I_RED + 1) {
      if constexpr (!EM) psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    } else if constexpr (w == I_RED + 2) {
      if constexpr (!EM) psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    } else if constexpr (w == I_INV) {
      if constexpr (!EM) pinv = safe_rcp(psum);
    } else if constexpr (w >= I_ST && w < I_LD) {   // posterior of slice t: normalise(alpha_t * beta_t)
      constexpr int n = w - I_ST;
      if constexpr (EM) {  // E-step: the carried beta_t itself
        if (on) reinterpret_cast<double2*>(rt + (row0 + t_cur) * SP)[4 * n + q] = make_double2(beta[n][0], beta[n][1]);
        return;
      }
      double* prow = post + (row0 + t_cur) * post_stride + post_off;
      if (VEC) {
        if (on) reinterpret_cast<double2*>(prow)[4 * n + q] = make_double2(a[n][0] * pinv, a[n][1] * pinv);
      } else {
        const int col = 8 * n + 2 * q;
        if (on && col < C.S) prow[col] = a[n][0] * pinv;
        if (on && col + 1 < C.S) prow[col + 1] = a[n][1] * pinv;
      }
    } else if constexpr (w >= I_LD && w < I_DOT) {
      if constexpr (!EM) {
        constexpr int n = w - I_LD;
        a[n][0] = an[n][0];
        a[n][1] = an[n][1];
      }
    } else if constexpr (w >= I_DOT && w < I_DRED) {
      constexpr int n = w - I_DOT;
      const double2 v = csv[4 * n + q];
      if constexpr (n == 0) { d0 = r[n][0] * v.x; d1 = r[n][1] * v.y; }
      else { d0 += r[n][0] * v.x; d1 += r[n][1] * v.y; }
    } else if constexpr (w == I_DRED) {
      dsum = d0 + d1;
    } else if constexpr (w == I_DRED + 1) {
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    } else if constexpr (w == I_DRED + 2) {
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
    } else if constexpr (w == I_H) {
      h = safe_rcp(dsum);
      hs = first_next ? 1.0 : h;
      if constexpr (EM) {
        if (on && q == 0) hvec[row0 + t_cur] = h;
      }
    } else if constexpr (w >= I_LAM && w < W_ITEMS) {
      constexpr int n = w - I_LAM;
      lam[n][0] *= hs;
      lam[n][1] *= hs;
    }
  };

  // lambda_{t-1} (used by the last side items of sweep t) and alpha_{t-1} (used at the start
  // of sweep t-1) are requested at the END of iteration t+1: ptxas sinks a load towards its
  // first use but not across the loop's back edge, so they are in flight for a whole sweep.
  {
    const bool p0 = Tw >= 2 && Tw - 2 < T;
    load_row(C.lam_comb + (long long)c_pre * SP, p0, lam);
    if (!EM) load_row(alpha + (row0 + Tw - 2) * AS, p0 && arow, an);
    if (Tw >= 3 && Tw - 3 < T) c_pre = __ldg(cfg + Tw - 3);
  }
  for (int t = Tw - 1; t >= 1; t--) {
    on = t < T;
    pre = t - 1 < T;                 // slice t-1 exists for this row
    first_next = t - 1 == T - 1;     // ... and is the row's last slice: beta = 1 there
    t_cur = t;
    // u = r . A^T  (k = current state, n = previous state)
    mma_sweep<NT>(u, r, frag, [&](auto sc) { run_items<2 * NT * NT, W_ITEMS, decltype(sc)::value>(item); });
#pragma unroll
    for (int n = 0; n < NT; n++) {
      beta[n][0] = first_next ? 1.0 : u[n][0] * h;
      beta[n][1] = first_next ? 1.0 : u[n][1] * h;
      r[n][0] = first_next ? lam[n][0] : u[n][0] * lam[n][0];
      r[n][1] = first_next ? lam[n][1] : u[n][1] * lam[n][1];
    }
    {  // requests for iteration t-1: lambda_{t-2}, alpha_{t-2}, evidence index of slice t-3
      const bool p2 = t >= 2 && t - 2 < T;
      load_row(C.lam_comb + (long long)c_pre * SP, p2, lam);
      if (!EM) load_row(alpha + (row0 + t - 2) * AS, p2 && arow, an);
      if (t >= 3 && t - 3 < T) c_pre = __ldg(cfg + t - 3);
    }
  }
  if (Tw >= 1) {   // posterior of slice 0
    on = 0 < T;
    t_cur = 0;
    static_for<0, I_LD>(item);
  }
  if (EM && Tw >= 1) {  // first slice: joint = A0 * r_0 / (phi0 . r_0)
    double z0 = 0, z1 = 0;
#pragma unroll
    for (int n = 0; n < NT; n++) {
      z0 += C.phi0[8 * n + 2 * q] * r[n][0];
      z1 += C.phi0[8 * n + 2 * q + 1] * r[n][1];
    }
    const double zinv = safe_rcp(quad_sum_full(z0 + z1));
    if (0 < T) {
      double2* out0 = reinterpret_cast<double2*>(r0 + (long long)orig * SP);
#pragma unroll
      for (int n = 0; n < NT; n++) out0[4 * n + q] = make_double2(r[n][0] * zinv, r[n][1] * zinv);
    }
  }
}

#include "chain_pair.cuh"
#include "chain_small.cuh"

// ----------------------------------------------------------------- EM ----
// Sufficient statistics of the whole batch in one pass over the two row stores
// (own[k] from the forward kernel, bt[k] = beta_k and hv[k] = h_k from the backward kernel):
//   N_k      = own_k . beta_k
//   G[i][j] += own_{k-1}[i] (h_k / N_{k-1}) * lambda_{c_k}[j] beta_k[j]   (rows that do not open a series)
//   Cc[c_k][j] += own_k[j] beta_k[j] / N_k                                (posterior of the row)
// G is a DMMA GEMM over the rows (split-K over CTAs, warp w owns the 8 previous states
// 8w..8w+7); the row-wise scalars and the evidence row are folded into its operands while the
// tile sits in shared memory, and the posterior rows are added into `phases` private tables
// (thread = column x phase, rows in ascending order: deterministic).  Tiles of 32 rows arrive
// through a two-stage cp.async pipeline; partial results are reduced afterwards in a fixed order.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(32 * NT * (NT >= 2 ? 2 : 1), 1)
    k_chain_stats(const double* __restrict__ own, const double* __restrict__ bt, const double* __restrict__ hv,
                  const int* __restrict__ cfg, const unsigned char* __restrict__ first,
                  const double* __restrict__ lam_comb, long long rows, int n_comb, int phases,
                  double* __restrict__ partG, double* __restrict__ partC,
                  const double* __restrict__ rnv, int own_stride) {
  // rnv[k] = 1 / (own_k . beta_k) when the producing kernels know it from their scale bookkeeping
  // (warp-pair kernels); nullptr: the dot products are formed here
  // NT m-tiles x GR halves of the n-tiles: 2 NT warps keep four warps on every scheduler
  constexpr int SP = 8 * NT, KC = 32, LD = SP + 4, GR = NT >= 2 ? 2 : 1, NW = NT * GR, NTH = 32 * NW;
  constexpr int TR = KC + 1, NH = NT / GR, STG = 3 * TR * LD;
  constexpr int RW = (TR + NW - 1) / NW, CW = (SP + 31) / 32;   // tile rows per warp, columns per lane
  extern __shared__ double sm[];
  // stage s (2): Aop rows at sm + s * STG, Bop rows TR * LD further, posterior rows 2 TR LD further;
  // tile row i of the tile starting at k0 holds data row k0 - 1 + i
  double* s_h = sm + 2 * STG;               // [3][TR + 1] h of the tile's rows (tile t in slot t % 3)
  double* s_rn = s_h + 3 * (TR + 1) + 1;    // [3][TR + 1] 1 / N of the tile's rows
  double* s_lam = s_rn + 3 * (TR + 1) + 1;  // [n_comb][SP] evidence rows
  const int tab = n_comb * SP;
  double* s_tab = s_lam + tab;              // [phases][n_comb][SP]
  int* s_c = reinterpret_cast<int*>(s_tab + (long long)phases * tab);  // [3][TR + 1] evidence index, -1 - c: opens a series
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, w = threadIdx.x >> 5;
  const int mt = w % NT, nh = w / NT;
  for (int x = threadIdx.x; x < phases * tab; x += NTH) s_tab[x] = 0.0;
  for (int x = threadIdx.x; x < tab; x += NTH) s_lam[x] = lam_comb[x];

  long long chunk = (rows + gridDim.x - 1) / gridDim.x;
  chunk = (chunk + KC - 1) / KC * KC;
  const long long k_begin = (long long)blockIdx.x * chunk;
  const long long k_end = k_begin + chunk < rows ? k_begin + chunk : rows;

  // The rows of the next tile travel through registers: requested before the tensor work of
  // the current tile, turned into operands after it (no staging copy in shared memory).
  double ro[RW][CW], rb[RW][CW];
  auto rows_load = [&](long long k0) {
#pragma unroll
    for (int u = 0; u < RW; u++) {
      const int i = w + u * NW;
      const long long k = k0 - 1 + i;
      const bool okr = i < TR && k >= 0 && k < k_end;
#pragma unroll
      for (int c = 0; c < CW; c++) {
        const int col = lane + 32 * c;
        const bool ok = okr && col < SP;
        ro[u][c] = (ok && col < own_stride) ? __ldg(own + k * own_stride + col) : 0.0;
        rb[u][c] = ok ? __ldg(bt + k * SP + col) : 0.0;
      }
    }
  };
  // per-row scalars run one tile further ahead (thread i <-> tile row i): row i needs row i+1's
  constexpr int MS = (TR + NTH - 1) / NTH;
  int m_c[MS];
  unsigned char m_f[MS];
  double m_h[MS], m_rn[MS];
  auto meta_load = [&](long long k0) {  // only requests: nothing here waits for the values
#pragma unroll
    for (int u = 0; u < MS; u++) {
      const int i = threadIdx.x + u * NTH;
      const long long k = k0 - 1 + i;
      const bool ok = i < TR && k >= 0 && k < k_end;
      m_f[u] = ok ? __ldg(first + k) : 0;
      m_c[u] = ok ? __ldg(cfg + k) : 0;
      m_h[u] = ok ? __ldg(hv + k) : 0.0;
      m_rn[u] = (ok && rnv) ? __ldg(rnv + k) : 0.0;
    }
  };
  auto meta_store = [&](int slot) {
#pragma unroll
    for (int u = 0; u < MS; u++) {
      const int i = threadIdx.x + u * NTH;
      if (i < TR) {
        s_c[slot * (TR + 1) + i] = m_f[u] ? -1 - m_c[u] : m_c[u];
        s_h[slot * (TR + 1) + i] = m_h[u];
        s_rn[slot * (TR + 1) + i] = m_rn[u];
      }
    }
  };
  // registers -> operands of the tile starting at k0, in stage `stage`; its scalars sit in `slot`
  auto rows_pass = [&](int stage, int slot, long long k0) {
    double* so = sm + stage * STG;
    double* sb = so + TR * LD;
    double* sp = sb + TR * LD;
    const int* sc = s_c + slot * (TR + 1);
    const double* sh = s_h + slot * (TR + 1);
    const double* srn = s_rn + slot * (TR + 1);
    double N[RW];
    if (!rnv) {
#pragma unroll
      for (int u = 0; u < RW; u++) {
        double part = 0;
#pragma unroll
        for (int c = 0; c < CW; c++) part += ro[u][c] * rb[u][c];
        N[u] = part;
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1)
#pragma unroll
        for (int u = 0; u < RW; u++) N[u] += __shfl_xor_sync(0xffffffffu, N[u], sft);
    }
#pragma unroll
    for (int u = 0; u < RW; u++) {
      const int i = w + u * NW;
      if (i < TR) {
        const double rn = rnv ? srn[i] : (N[u] != 0 ? 1.0 / N[u] : 0.0);
        // the pair (this row, next row) counts unless the next row opens a series or lies outside
        const bool pair = i < KC && k0 + i < k_end && sc[i + 1] >= 0;
        const double wk = pair ? sh[i + 1] * rn : 0.0;
        const int cc = sc[i] >= 0 ? sc[i] : -1 - sc[i];
        const double* lam = s_lam + cc * SP;
#pragma unroll
        for (int c = 0; c < CW; c++) {
          const int col = lane + 32 * c;
          if (col < SP) {
            sp[i * LD + col] = ro[u][c] * rb[u][c] * rn;
            so[i * LD + col] = ro[u][c] * wk;
            sb[i * LD + col] = rb[u][c] * lam[col];
          }
        }
      }
    }
  };

  double acc[NH > 0 ? NH : 1][2];
#pragma unroll
  for (int n = 0; n < NH; n++) acc[n][0] = acc[n][1] = 0.0;
  if (k_begin < k_end) {  // prologue: tile 0 becomes operands, the scalars of tile 1 are parked
    meta_load(k_begin);
    meta_store(0);
    rows_load(k_begin);
    __syncthreads();
    rows_pass(0, 0, k_begin);
    if (k_begin + KC < k_end) {
      meta_load(k_begin + KC);
      meta_store(1);
    }
  }
  __syncthreads();
  int stage = 0, slot = 0;
  for (long long k0 = k_begin; k0 < k_end; k0 += KC, stage ^= 1, slot = slot == 2 ? 0 : slot + 1) {
    const int slot1 = slot == 2 ? 0 : slot + 1, slot2 = slot1 == 2 ? 0 : slot1 + 1;
    const bool more = k0 + KC < k_end, more2 = k0 + 2 * KC < k_end;
    if (more) rows_load(k0 + KC);
    if (more2) meta_load(k0 + 2 * KC);
    const double* so = sm + stage * STG;
    const double* sb = so + TR * LD;
    const double* sp = sb + TR * LD;
    const int* sc = s_c + slot * (TR + 1);
    // ---- G += Aop^T Bop (rows k0 .. k0+KC-1: A from tile row i-1, B from tile row i) ----
    if constexpr (NT == 8) {
      // 2 x 2 tiles per warp: two A and two B fragments feed four tensor instructions
      const int m2 = 2 * (w & 3), n2 = 2 * (w >> 2);
#pragma unroll
      for (int kk = 0; kk < KC / 4; kk++) {
        const double a0 = so[(4 * kk + q) * LD + 8 * m2 + g], a1 = so[(4 * kk + q) * LD + 8 * (m2 + 1) + g];
        const double b0 = sb[(4 * kk + q + 1) * LD + 8 * n2 + g], b1 = sb[(4 * kk + q + 1) * LD + 8 * (n2 + 1) + g];
        dmma(acc[0][0], acc[0][1], a0, b0);
        dmma(acc[1][0], acc[1][1], a0, b1);
        dmma(acc[2][0], acc[2][1], a1, b0);
        dmma(acc[3][0], acc[3][1], a1, b1);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KC / 4; kk++) {
        const double a = so[(4 * kk + q) * LD + 8 * mt + g];
#pragma unroll
        for (int n = 0; n < NH; n++)
          dmma(acc[n][0], acc[n][1], a, sb[(4 * kk + q + 1) * LD + 8 * (nh * NH + n) + g]);
      }
    }
    // ---- posterior rows into the evidence-indexed tables ----
    {
      const int col = threadIdx.x % SP, ph = threadIdx.x / SP;
      if (ph < phases) {
        double* mine = s_tab + (long long)ph * tab + col;
        for (int i = 1 + ph; i < TR; i += phases) {
          const int cc = sc[i] >= 0 ? sc[i] : -1 - sc[i];
          mine[cc * SP] += sp[i * LD + col];
        }
      }
    }
    // ---- the next tile: registers -> operands in the other stage ----
    if (more) rows_pass(stage ^ 1, slot1, k0 + KC);
    if (more2) meta_store(slot2);  // scalars of tile k0 + 2 KC: a slot nobody reads in this iteration
    __syncthreads();
  }
  double* outG = partG + (long long)blockIdx.x * SP * SP;
#pragma unroll
  for (int n = 0; n < NH; n++) {
    int mi = mt, ni = nh * NH + n;
    if (NT == 8) { mi = 2 * (w & 3) + (n >> 1); ni = 2 * (w >> 2) + (n & 1); }
    outG[(8 * mi + g) * SP + 8 * ni + 2 * q] = acc[n][0];
    outG[(8 * mi + g) * SP + 8 * ni + 2 * q + 1] = acc[n][1];
  }
  double* outC = partC + (long long)blockIdx.x * tab;
  for (int x = threadIdx.x; x < tab; x += NTH) {
    double s = 0;
    for (int ph = 0; ph < phases; ph++) s += s_tab[(long long)ph * tab + x];
    outC[x] = s;
  }
}

// first[k] = 1 on the rows that open a series
__global__ void k_chain_first(const long long* row_off, int n_series, long long rows, unsigned char* first) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_series && row_off[s] < rows) first[row_off[s]] = 1;
}

// out[x] = sum_p part[p][x] in a fixed association: a CTA owns 32 outputs, warp g adds the parts
// g, g+8, g+16, ... in order (two running sums: the loads of a thread do not wait for each other),
// the eight group sums are added in order.  One thread per output walking all the parts was a
// chain of `parts` dependent loads (37 us for 148 parts of a 64 x 64 table).
__global__ void __launch_bounds__(256) k_chain_sum_parts(const double* __restrict__ part, int parts, long long n,
                                                         double* __restrict__ out) {
  __shared__ double red[8][33];
  const int lx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long long x = blockIdx.x * 32LL + lx;
  double s0 = 0, s1 = 0;
  if (x < n) {
    int p = g;
    for (; p + 8 < parts; p += 16) {
      s0 += part[(long long)p * n + x];
      s1 += part[(long long)(p + 8) * n + x];
    }
    if (p < parts) s0 += part[(long long)p * n + x];
  }
  red[g][lx] = s0 + s1;
  __syncthreads();
  if (g == 0 && x < n) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += red[k][lx];
    out[x] = s;
  }
}

// column sums of r0 [n_series][SP], first stage: CTA c adds the rows [c * per, (c + 1) * per) — whole
// rows per warp-load, 256 / SP rows at a time — into part[c][SP] (fixed order); k_chain_sum_parts
// adds the CTAs' rows.  (One block per column read one double per 512-byte row.)
__global__ void __launch_bounds__(256) k_chain_g0_part(const double* __restrict__ r0, int n_series, int SP, int per,
                                                       double* __restrict__ part) {
  __shared__ double red[256];
  const int b0 = blockIdx.x * per, b1 = min(n_series, b0 + per);
  if (SP > 256 || 256 % SP != 0) {   // wide interfaces (dense engine): a thread per column, rows in order
    for (int col = threadIdx.x; col < SP; col += 256) {
      double s = 0;
      for (int b = b0; b < b1; b++) s += r0[(long long)b * SP + col];
      part[(long long)blockIdx.x * SP + col] = s;
    }
    return;
  }
  const int col = threadIdx.x % SP, grp = threadIdx.x / SP, groups = 256 / SP;
  double s = 0;
  for (int b = b0 + grp; b < b1; b += groups) s += r0[(long long)b * SP + col];
  red[threadIdx.x] = s;
  __syncthreads();
  if (grp == 0) {
    double t = 0;
    for (int k = 0; k < groups; k++) t += red[k * SP + col];
    part[(long long)blockIdx.x * SP + col] = t;
  }
}

// expected table of the interface clique over all slices (E) and over first slices (E0)
__global__ void k_chain_expect_c0(const double* base0, const double* base1, const int* ent_im,
                                  const int* ent_ip, int n, int SP, const double* G,
                                  const double* g0, double* E, double* E0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double first = base0[e] * g0[ent_ip[e]];
  E0[e] = first;
  E[e] = base1[e] * G[ent_im[e] * SP + ent_ip[e]] + first;
}

struct LeafExpect {
  int n_free, card[8], stride[8];   // free variables of the leaf, their cfg strides
  int slot;                          // position among the plan's active leaves, -1 = no evidence columns
  int mult, n_cfg, miss_cfg;         // combined-index digit of this leaf
  int m, R, S, SP, n_comb;
  long long lam_off;
};

// expected table of a leaf clique: E(s, y) = base(s, y) * sum over evidence configurations
// compatible with y of  [posterior mass of s observed under that configuration] / Lambda(cfg, s)
__global__ void k_chain_expect_leaf(const double* base1, const int* pbase, const int* poff,
                                    const int* ip_to_s, LeafExpect L, const double* Cc,
                                    const double* lam, double* E) {
  // a warp per table entry: the lanes share the sum over the evidence configurations
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (x >= L.m * L.R) return;
  const int s = x / L.R, r = x - s * L.R;
  const int entry = pbase[s] + poff[r];
  int digit[8], rem = r;
  for (int k = 0; k < L.n_free; k++) { digit[k] = rem % L.card[k]; rem /= L.card[k]; }
  double W = 0;
  for (int mask = 0; mask < (1 << L.n_free); mask++) {   // each free variable: observed as y_k, or not observed
    int cl = 0;
    for (int k = 0; k < L.n_free; k++) cl += ((mask >> k) & 1 ? L.card[k] : digit[k]) * L.stride[k];
    if (L.slot < 0 && cl != L.miss_cfg) continue;
    double mass = 0, lam_s = 0;
    for (int ip = 0; ip < L.S; ip++) {
      if (ip_to_s[ip] != s) continue;
      lam_s = lam[L.lam_off + (long long)cl * L.SP + ip];
      for (int c = lane; c < L.n_comb; c += 32)
        if (L.slot < 0 || (c / L.mult) % L.n_cfg == cl) mass += Cc[(long long)c * L.SP + ip];
    }
    mass = warp_sum(mass);
    if (lam_s != 0) W += mass / lam_s;
  }
  if (lane == 0) E[entry] = base1[entry] * W;
}

// counts[j] = pseudo + sum_r E[base[j] + off[r]]   (family table of one variable)
__global__ void k_chain_family(const double* E, const int* pbase, const int* poff, int m, int R,
                               double pseudo, double* counts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s = 0;
  const int b = pbase[j];
  for (int r = 0; r < R; r++) s += E[b + poff[r]];
  counts[j] = pseudo + s;
}

__global__ void k_chain_tail(const double* ll, const int* status, int n_series, double* tail) {
  __shared__ double red[40];
  double L = 0, bad = 0;
  for (int i = threadIdx.x; i < n_series; i += blockDim.x) { L += ll[i]; bad += status[i] ? 1.0 : 0.0; }
  L = block_sum(L, red);
  bad = block_sum(bad, red);
  if (threadIdx.x == 0) { tail[0] = L; tail[1] = bad != 0 ? 1.0 : 0.0; }
}

template <class K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    NIPGPU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NIPGPU_OK;
}

// zc[series] = sum of the series' last forward row, zf[series] = the exponent that row carries:
// Z = zc 2^-zf is the series' evidence mass in the units of the bookkeeping (one warp per series)
__global__ void k_chain_final(const double* __restrict__ alpha, const int* __restrict__ fexp,
                              const long long* __restrict__ row_off, const int* __restrict__ order,
                              const int* __restrict__ len_sorted, int n_series, int SP, double* zc, int* zf) {
  const int bp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (bp >= n_series) return;
  const int T = len_sorted[bp], orig = order[bp];
  double s = 0;
  if (T > 0) {
    const double* row = alpha + (row_off[orig] + T - 1) * SP;
    for (int j = lane; j < SP; j += 32) s += row[j];
  }
  s = warp_sum(s);
  if (lane == 0) {
    zc[orig] = s;
    zf[orig] = T > 0 ? fexp[row_off[orig] + T - 1] : 0;
  }
}

#ifndef NIPGPU_TEAM4_MAX_SERIES
#define NIPGPU_TEAM4_MAX_SERIES 3072
#endif
// NT = 4 / 8: every 8-sequence group on a team of warps (chain_pair.cuh): two warps of one
// scheduler, or (8 state tiles) four warps on the SM's four schedulers.  NIPGPU_CHAIN_PAIR=0
// keeps the one-warp kernels, NIPGPU_CHAIN_TEAM=2|4 forces the team width (A/B timing).
static int team_width(int NT, int n_series) {
  static const int forced = [] {
    const char* p = getenv("NIPGPU_CHAIN_PAIR");
    if (p && p[0] == '0') return 1;
    const char* e = getenv("NIPGPU_CHAIN_TEAM");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || (NT != 4 && NT != 8)) return 1;
  if (NT == 4) return 2;
  if (forced == 2 || forced == 4) return forced;
  // measured on C2's model (profiles/r02_team_width.txt): four warps on four schedulers take
  // 1.50 / 1.55 / 2.01 / 3.40 ms per smoothing pass of 512 / 1024 / 2048 / 4096 sequences, two
  // warps on one scheduler 2.79 ms for any batch up to one group per scheduler (4736 sequences)
  return n_series <= NIPGPU_TEAM4_MAX_SERIES ? 4 : 2;
}
template <int NT, int W>
static int team_grid(int n_series) {
  const int groups = (n_series + 7) / 8;
  return (groups + TeamGeom<NT, W>::TEAMS - 1) / TeamGeom<NT, W>::TEAMS;
}

// runs f(std::integral_constant<int, W>) for the team width chosen for NT; false: one-warp kernels
template <int NT, class F>
static bool with_team(int n_series, F&& f, int* err) {
  if constexpr (NT == 4 || NT == 8) {
    const int w = team_width(NT, n_series);
    if (w == 2) { *err = f(std::integral_constant<int, 2>{}); return true; }
    if constexpr (NT == 8) {
      if (w == 4) { *err = f(std::integral_constant<int, 4>{}); return true; }
    }
  }
  return false;
}

template <int NT, bool FILT, bool WLL>
int launch_forward_v(const ChainDev& C, const ChainBatchDev& B, const ChainInferArgs& a, double* alpha,
                     cudaStream_t st) {
  {
    int err = NIPGPU_OK;
    if (with_team<NT>(B.n_series, [&](auto wc) -> int {
          constexpr int W = decltype(wc)::value;
          using G = TeamGeom<NT, W>;
          const size_t smem = G::smem_bytes();
          if (int e = set_smem(k_chain_forward_team<NT, W, FILT, WLL>, smem)) return e;
          k_chain_forward_team<NT, W, FILT, WLL><<<team_grid<NT, W>(B.n_series), G::THREADS, smem, st>>>(
              C, B, alpha, a.d_post, a.post_stride, a.post_off, a.d_ll, a.d_status);
          NIPGPU_LAUNCHED();
          if (!FILT && B.n_series > 0) {   // a backward pass follows: its normalisers
            k_chain_final<<<(B.n_series + 7) / 8, 256, 0, st>>>(alpha, B.fexp, B.row_off, B.order, B.len_sorted,
                                                              B.n_series, C.SP, B.zc, B.zf);
            NIPGPU_LAUNCHED();
          }
          return NIPGPU_OK;
        }, &err))
      return err;
  }
  const size_t smem = sizeof(double) * (64 * NT * NT + 16 * NT);
  if (int e = set_smem(k_chain_forward<NT, FILT, WLL>, smem)) return e;
  const int grid = (B.n_series + 31) / 32;
  k_chain_forward<NT, FILT, WLL><<<grid, 128, smem, st>>>(C, B, alpha, a.d_post, a.post_stride,
                                                          a.post_off, a.d_ll, a.d_status);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

template <int NT>
int launch_forward(const ChainDev& C, const ChainBatchDev& B, const ChainInferArgs& a, double* alpha,
                   cudaStream_t st) {
  const bool filt = a.forward_only && a.d_post;
  if (filt) return a.want_ll ? launch_forward_v<NT, true, true>(C, B, a, alpha, st)
                             : launch_forward_v<NT, true, false>(C, B, a, alpha, st);
  return a.want_ll ? launch_forward_v<NT, false, true>(C, B, a, alpha, st)
                   : launch_forward_v<NT, false, false>(C, B, a, alpha, st);
}

template <int NT>
int launch_backward(const ChainDev& C, const ChainBatchDev& B, const ChainInferArgs& a,
                    const double* alpha, cudaStream_t st) {
  const int grid = (B.n_series + 31) / 32;
  const size_t smem = sizeof(double) * (64 * NT * NT + 16 * NT);
  const bool vec = ((a.post_stride | a.post_off) & 1) == 0 && C.S == C.SP;
  {
    int err = NIPGPU_OK;
    if (with_team<NT>(B.n_series, [&](auto wc) -> int {
          constexpr int W = decltype(wc)::value;
          using G = TeamGeom<NT, W>;
          const size_t psm = G::smem_bytes();
          const int tg = team_grid<NT, W>(B.n_series);
          if (vec) {
            if (int e = set_smem(k_chain_backward_team<NT, W, true, false>, psm)) return e;
            k_chain_backward_team<NT, W, true, false><<<tg, G::THREADS, psm, st>>>(C, B, alpha, a.d_post, a.post_stride,
                                                                                   a.post_off, nullptr, nullptr, nullptr);
          } else {
            if (int e = set_smem(k_chain_backward_team<NT, W, false, false>, psm)) return e;
            k_chain_backward_team<NT, W, false, false><<<tg, G::THREADS, psm, st>>>(C, B, alpha, a.d_post, a.post_stride,
                                                                                    a.post_off, nullptr, nullptr, nullptr);
          }
          NIPGPU_LAUNCHED();
          return NIPGPU_OK;
        }, &err))
      return err;
  }
  if (vec) {
    if (int e = set_smem(k_chain_backward<NT, true, false>, smem)) return e;
    k_chain_backward<NT, true, false><<<grid, 128, smem, st>>>(C, B, alpha, a.d_post, a.post_stride,
                                                                a.post_off, nullptr, nullptr, nullptr);
  } else {
    if (int e = set_smem(k_chain_backward<NT, false, false>, smem)) return e;
    k_chain_backward<NT, false, false><<<grid, 128, smem, st>>>(C, B, alpha, a.d_post, a.post_stride,
                                                                 a.post_off, nullptr, nullptr, nullptr);
  }
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

struct ChainStatsArgs {
  double *bt, *hv, *r0;
  const unsigned char* first;
  long long rows;
  int n_comb, phases, parts;
  size_t smem;
  double *partG, *partC;
};

template <int NT>
int launch_em(const ChainDev& C, const ChainBatchDev& B, const ChainInferArgs& a, double* alpha,
              const ChainStatsArgs& s, cudaStream_t st) {
  if (int e = launch_forward_v<NT, false, true>(C, B, a, alpha, st)) return e;
  const int grid = (B.n_series + 31) / 32;
  const size_t smem = sizeof(double) * (64 * NT * NT + 16 * NT);
  int err = NIPGPU_OK;
  const bool paired = with_team<NT>(B.n_series, [&](auto wc) -> int {
    constexpr int W = decltype(wc)::value;
    using G = TeamGeom<NT, W>;
    const size_t psm = G::smem_bytes();
    if (int e = set_smem(k_chain_backward_team<NT, W, true, true>, psm)) return e;
    k_chain_backward_team<NT, W, true, true><<<team_grid<NT, W>(B.n_series), G::THREADS, psm, st>>>(
        C, B, alpha, nullptr, C.SP, 0, s.bt, s.r0, s.hv);
    return NIPGPU_OK;
  }, &err);
  if (err) return err;
  if (!paired) {
    if (int e = set_smem(k_chain_backward<NT, true, true>, smem)) return e;
    k_chain_backward<NT, true, true><<<grid, 128, smem, st>>>(C, B, alpha, nullptr, C.SP, 0, s.bt, s.r0, s.hv);
  }
  NIPGPU_LAUNCHED();
  if (int e = set_smem(k_chain_stats<NT>, s.smem)) return e;
  k_chain_stats<NT><<<s.parts, 32 * NT * (NT >= 2 ? 2 : 1), s.smem, st>>>(alpha, s.bt, s.hv, B.cfg, s.first, C.lam_comb, s.rows,
                                                      s.n_comb, s.phases, s.partG, s.partC,
                                                      paired ? B.rn_out : nullptr, NT == 1 ? C.AS : C.SP);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

template <class T>
int upload(T** dst, const std::vector<T>& src, cudaStream_t st) {
  const size_t n = std::max<size_t>(src.size(), 1);
  NIPGPU_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
  if (!src.empty())
    NIPGPU_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return NIPGPU_OK;
}

}  // namespace

// ------------------------------------------------------------------ host ---
std::string chain_build(HostModel& hm, ChainModel& cm) {
  cm.ok = false;
  if (!hm.chain_ok) return hm.chain_why;
  const int S = hm.S;
  int SP = 8;
  while (SP < S) SP *= 2;
  cm.dense = SP > 64;          // large interface: per-slice tiled GEMM (dense.cu)
  if (cm.dense) SP = (S + 127) / 128 * 128;
  if (SP > 4096) return "interface larger than 4096 states";
  cm.S = S; cm.SP = SP; cm.NT = cm.dense ? 0 : SP / 8; cm.c0 = hm.in_clique;
  const int c0 = cm.c0, nd = hm.clique_dim(c0);
  // entry -> (previous-slice interface state, current interface state)
  std::vector<int> istride_prev(hm.nv, 0), istride_cur(hm.nv, 0);
  int st = 1;
  for (int k = 0; k < hm.nif; k++) {
    istride_prev[hm.prev[k]] = st;
    istride_cur[hm.outg[k]] = st;
    st *= hm.card[hm.outg[k]];
  }
  cm.ent_of.assign((size_t)S * S, 0);
  for (int ent = 0; ent < hm.csize[c0]; ent++) {
    int rem = ent, im = 0, ip = 0;
    for (int k = 0; k < nd; k++) {
      const int v = hm.clique_vars(c0)[k], digit = rem % hm.card[v];
      rem /= hm.card[v];
      bool is_prev = false;
      for (int x = 0; x < hm.nif; x++)
        if (hm.prev[x] == v) is_prev = true;
      if (is_prev) im += digit * istride_prev[v];
      else ip += digit * istride_cur[v];
    }
    cm.ent_of[(size_t)im * S + ip] = ent;
  }
  cm.var_leaf.assign(hm.nv, -1);
  cm.var_slot.assign(hm.nv, -1);
  cm.leaves.clear();
  long long lam_total = 0;
  for (size_t li = 0; li < hm.leaves.size(); li++) {
    ChainLeafHost L;
    L.clique = hm.leaves[li];
    const int s = hm.leaf_sepset[li];
    std::vector<int> sv(hm.sepset_vars(s), hm.sepset_vars(s) + hm.sepset_dim(s));
    long long ncfg = 1;
    for (int k = 0; k < hm.clique_dim(L.clique); k++) {
      const int v = hm.clique_vars(L.clique)[k];
      if (std::find(sv.begin(), sv.end(), v) != sv.end()) continue;
      L.free_vars.push_back(v);
      L.cfg_stride.push_back((int)ncfg);
      ncfg *= hm.card[v] + 1;
      if (ncfg * SP > (1LL << 27)) return "leaf clique with too many evidence configurations";
    }
    L.n_cfg = (int)ncfg;
    L.miss_cfg = 0;
    for (size_t k = 0; k < L.free_vars.size(); k++) L.miss_cfg += hm.card[L.free_vars[k]] * L.cfg_stride[k];
    L.proj = hm.add_proj(L.clique, sv);
    L.ip_to_s.assign(S, 0);
    for (int ip = 0; ip < S; ip++) {
      int sidx = 0, sst = 1;
      for (int v : sv) {
        const int digit = (ip / istride_cur[v]) % hm.card[v];
        sidx += digit * sst;
        sst *= hm.card[v];
      }
      L.ip_to_s[ip] = sidx;
    }
    L.lam_off = lam_total;
    lam_total += (long long)L.n_cfg * SP;
    for (size_t k = 0; k < L.free_vars.size(); k++) {
      cm.var_leaf[L.free_vars[k]] = (int)cm.leaves.size();
      cm.var_slot[L.free_vars[k]] = (int)k;
    }
    cm.leaves.push_back(L);
  }
  cm.n_real = (int)cm.leaves.size();
  for (int k = 0; k < hm.nif; k++) {  // evidence on an I_t variable itself: indicator "leaf"
    ChainLeafHost L;
    L.var = hm.outg[k];
    L.free_vars = {L.var};
    L.cfg_stride = {1};
    L.n_cfg = hm.card[L.var] + 1;
    L.miss_cfg = hm.card[L.var];
    L.ip_to_s.assign(S, 0);
    for (int ip = 0; ip < S; ip++) L.ip_to_s[ip] = (ip / istride_cur[L.var]) % hm.card[L.var];
    L.lam_off = lam_total;
    lam_total += (long long)L.n_cfg * SP;
    cm.var_leaf[L.var] = (int)cm.leaves.size();
    cm.var_slot[L.var] = 0;
    cm.leaves.push_back(L);
  }
  cm.lam_total = lam_total;
  cm.ok = true;
  return "";
}

int chain_upload_structure(const HostModel& hm, ChainModel& cm, cudaStream_t st) {
  if (!cm.ok) return NIPGPU_OK;
  const size_t sp2 = (size_t)cm.SP * cm.SP;
  if (int e = upload(&cm.d_ent_of, cm.ent_of, st)) return e;
  {  // entry of the interface clique -> (previous state, current state); leaves: state -> sepset entry
    std::vector<int> im((size_t)hm.csize[cm.c0]), ip((size_t)hm.csize[cm.c0]), i2s;
    for (int a = 0; a < cm.S; a++)
      for (int b = 0; b < cm.S; b++) { im[cm.ent_of[(size_t)a * cm.S + b]] = a; ip[cm.ent_of[(size_t)a * cm.S + b]] = b; }
    for (int l = 0; l < cm.n_real; l++) i2s.insert(i2s.end(), cm.leaves[l].ip_to_s.begin(), cm.leaves[l].ip_to_s.end());
    if (int e = upload(&cm.d_ent_im, im, st)) return e;
    if (int e = upload(&cm.d_ent_ip, ip, st)) return e;
    if (int e = upload(&cm.d_ip_to_s, i2s, st)) return e;
    NIPGPU_CUDA(cudaStreamSynchronize(st));
  }
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_Bf1, sp2 * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_Bb1, sp2 * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_Bb0, sp2 * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_phi0, cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_lam0, cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_R1, cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_colsum, cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cm.d_lam, std::max<long long>(cm.lam_total, 1) * sizeof(double)));
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_lam, 0, std::max<long long>(cm.lam_total, 1) * sizeof(double), st));
  // pseudo leaves never change: Lambda[o][ip] = [state(ip) == o], Lambda[card][ip] = 1
  for (size_t l = cm.n_real; l < cm.leaves.size(); l++) {
    const ChainLeafHost& L = cm.leaves[l];
    std::vector<double> tab((size_t)L.n_cfg * cm.SP, 0.0);
    for (int c = 0; c < L.n_cfg; c++)
      for (int ip = 0; ip < cm.S; ip++)
        tab[(size_t)c * cm.SP + ip] = (c == L.miss_cfg || L.ip_to_s[ip] == c) ? 1.0 : 0.0;
    NIPGPU_CUDA(cudaMemcpyAsync(cm.d_lam + L.lam_off, tab.data(), tab.size() * sizeof(double),
                                cudaMemcpyHostToDevice, st));
    NIPGPU_CUDA(cudaStreamSynchronize(st));
  }
  {  // per-leaf metadata of the refresh kernels, uploaded once
    std::vector<int> meta;
    std::vector<long long> miss_rows;
    cm.leaf_meta_off.clear();
    for (int l = 0; l < cm.n_real; l++) {
      const ChainLeafHost& L = cm.leaves[l];
      cm.leaf_meta_off.push_back((int)meta.size());
      meta.push_back((int)L.free_vars.size());
      for (int v : L.free_vars) meta.push_back(hm.card[v]);
      for (int sstride : L.cfg_stride) meta.push_back(sstride);
      miss_rows.push_back(L.lam_off + (long long)L.miss_cfg * cm.SP);
    }
    if (int e = upload(&cm.d_leaf_meta, meta, st)) return e;
    if (int e = upload(&cm.d_miss_rows, miss_rows, st)) return e;
    NIPGPU_CUDA(cudaStreamSynchronize(st));
  }
  return NIPGPU_OK;
}

int chain_refresh(const HostModel& hm, ChainModel& cm, const double* d_base0, const double* d_base1,
                  const std::vector<int>& tab_off, const int* d_ipool, const double* d_R1,
                  const double* d_m10, cudaStream_t st) {
  if (!cm.ok) return NIPGPU_OK;
  // mass of the evidence-free slice per previous interface state / of the first slice
  // (k_jt_calibrate): m1_t = alpha_{t-1} . R1
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_R1, 0, cm.SP * sizeof(double), st));
  NIPGPU_CUDA(cudaMemcpyAsync(cm.d_R1, d_R1, cm.S * sizeof(double), cudaMemcpyDeviceToDevice, st));
  NIPGPU_CUDA(cudaMemcpyAsync(&cm.m1_0, d_m10, sizeof(double), cudaMemcpyDeviceToHost, st));
  const size_t sp2 = (size_t)cm.SP * cm.SP;
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_Bf1, 0, sp2 * sizeof(double), st));
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_Bb1, 0, sp2 * sizeof(double), st));
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_Bb0, 0, sp2 * sizeof(double), st));
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_phi0, 0, cm.SP * sizeof(double), st));
  const int S = cm.S;
  if (cm.dense) {
    if (int e = dense_refresh_mats(cm, d_base1 + tab_off[cm.c0], st)) return e;
  } else {
    k_chain_mats<<<(S * S + 255) / 256, 256, 0, st>>>(d_base0 + tab_off[cm.c0], d_base1 + tab_off[cm.c0],
                                                      cm.d_ent_of, S, cm.NT, cm.d_Bf1, cm.d_Bb1, cm.d_Bb0);
    NIPGPU_LAUNCHED();
    if (cm.NT == 1) {
      if (!cm.d_As) NIPGPU_CUDA(cudaMalloc((void**)&cm.d_As, 64 * sizeof(double)));
      k_chain_small_A<<<1, 64, 0, st>>>(cm.d_Bf1, S, cm.d_As);
      NIPGPU_LAUNCHED();
    }
  }
  k_chain_phi0<<<(S + 3) / 4, 128, 0, st>>>(d_base0 + tab_off[cm.c0], cm.d_ent_of, S, cm.d_phi0);
  NIPGPU_LAUNCHED();
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_colsum, 0, cm.SP * sizeof(double), st));
  k_chain_phi0<<<(S + 3) / 4, 128, 0, st>>>(d_base1 + tab_off[cm.c0], cm.d_ent_of, S, cm.d_colsum);
  NIPGPU_LAUNCHED();
  for (int l = 0; l < cm.n_real; l++) {
    const ChainLeafHost& L = cm.leaves[l];
    const Proj& p = hm.projs[L.proj];
    const int n = L.n_cfg * S;
    k_chain_lambda<<<(n + 127) / 128, 128, 0, st>>>(
        d_base1 + tab_off[L.clique], d_ipool + p.base_pos, d_ipool + p.off_pos, p.R,
        cm.d_ip_to_s + (size_t)l * S, cm.d_leaf_meta + cm.leaf_meta_off[l], L.n_cfg, S, cm.SP,
        cm.d_lam + L.lam_off);
    NIPGPU_LAUNCHED();
  }
  k_chain_lam_prod<<<(cm.SP + 127) / 128, 128, 0, st>>>(cm.d_lam, cm.d_miss_rows, cm.n_real, S, cm.SP, cm.d_lam0);
  NIPGPU_LAUNCHED();
  NIPGPU_CUDA(cudaStreamSynchronize(st));   // m1_0 has landed on the host
  cm.param_version++;
  return NIPGPU_OK;
}

void chain_free(ChainModel& cm) {
  cudaFree(cm.d_ent_of); cudaFree(cm.d_Bf1); cudaFree(cm.d_Bb1); cudaFree(cm.d_Bb0);
  cudaFree(cm.d_phi0); cudaFree(cm.d_lam0); cudaFree(cm.d_lam); cudaFree(cm.d_leaf_meta); cudaFree(cm.d_miss_rows); cudaFree(cm.d_R1); cudaFree(cm.d_colsum); cudaFree(cm.d_ent_im); cudaFree(cm.d_ent_ip); cudaFree(cm.d_ip_to_s); cudaFree(cm.d_As);
  cm = ChainModel();
}

bool chain_plan(const HostModel& hm, const ChainModel& cm, int n_obs, const int* obs_vars,
                const uint8_t* use_evidence, ChainPlan& plan) {
  plan = ChainPlan();
  if (!cm.ok) return false;
  plan.col_leaf_slot.assign(n_obs, -1);
  plan.col_stride.assign(n_obs, 0);
  plan.col_card.assign(n_obs, 0);
  plan.col_mult.assign(n_obs, 0);
  for (int k = 0; k < n_obs; k++) {
    const int v = obs_vars[k];
    if (use_evidence && !use_evidence[v]) continue;
    const int l = cm.var_leaf[v];
    if (l < 0) return false;  // evidence on a previous-slice variable: generic engine
    int a = (int)(std::find(plan.active_leaf.begin(), plan.active_leaf.end(), l) - plan.active_leaf.begin());
    if (a == (int)plan.active_leaf.size()) plan.active_leaf.push_back(l);
    plan.col_leaf_slot[k] = a;
    plan.col_stride[k] = cm.leaves[l].cfg_stride[cm.var_slot[v]];
    plan.col_card[k] = hm.card[v];
  }
  plan.n_active = (int)plan.active_leaf.size();
  if (plan.n_active > 8) return false;
  long long mult = 1;
  plan.c_miss = 0;
  for (int a = 0; a < plan.n_active; a++) {
    const ChainLeafHost& L = cm.leaves[plan.active_leaf[a]];
    plan.mult.push_back((int)mult);
    plan.c_miss += L.miss_cfg * (int)mult;
    mult *= L.n_cfg;
    if (mult * cm.SP > (1LL << 25)) return false;  // combined evidence table too large
  }
  plan.n_comb = (int)mult;
  for (int k = 0; k < n_obs; k++)
    if (plan.col_leaf_slot[k] >= 0) plan.col_mult[k] = plan.mult[plan.col_leaf_slot[k]];
  return true;
}

int chain_batch_prepare(const ChainModel& cm, ChainBatch& cb, int n_series, const int* len,
                        long long rows, int t_max, cudaStream_t st) {
  (void)t_max;
  if (cb.ready) return NIPGPU_OK;
  cb.order.resize(n_series);
  for (int i = 0; i < n_series; i++) cb.order[i] = i;
  std::stable_sort(cb.order.begin(), cb.order.end(), [&](int a, int b) { return len[a] > len[b]; });
  cb.len_sorted.resize(n_series);
  for (int i = 0; i < n_series; i++) cb.len_sorted[i] = len[cb.order[i]];
  if (int e = upload(&cb.d_order, cb.order, st)) return e;
  if (int e = upload(&cb.d_len_sorted, cb.len_sorted, st)) return e;
  if (cm.NT == 1 && !cm.dense) {   // time-major offsets of the thread-per-sequence kernels
    const int tm = n_series > 0 ? std::max(cb.len_sorted[0], 0) : 0;
    std::vector<long long> toff(tm + 1, 0);
    int alive = n_series;
    for (int t = 0; t < tm; t++) {
      while (alive > 0 && cb.len_sorted[alive - 1] <= t) alive--;
      toff[t + 1] = toff[t] + alive;
    }
    if (int e = upload(&cb.d_toff, toff, st)) return e;
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_cfgT, std::max<long long>(rows, 1) * sizeof(int)));
    NIPGPU_CUDA(cudaStreamSynchronize(st));   // `toff` dies here
  }
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_alpha, std::max<long long>(rows, 1) * cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_fexp, std::max<long long>(rows, 1) * sizeof(int)));
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_zc, (size_t)std::max(n_series, 1) * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_zf, (size_t)std::max(n_series, 1) * sizeof(int)));
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_cfg, std::max<long long>(rows, 1) * sizeof(int)));
  NIPGPU_CUDA(cudaMalloc((void**)&cb.d_lam_static, cm.SP * sizeof(double)));
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  cb.ready = true;
  return NIPGPU_OK;
}

void chain_batch_free(ChainBatch& cb) {
  cudaFree(cb.d_order); cudaFree(cb.d_len_sorted); cudaFree(cb.d_cfg); cudaFree(cb.d_alpha);
  cudaFree(cb.d_fexp); cudaFree(cb.d_zc); cudaFree(cb.d_zf); cudaFree(cb.d_rn);
  cudaFree(cb.d_lam_static); cudaFree(cb.d_cols); cudaFree(cb.d_rows); cudaFree(cb.d_comb);
  cudaFree(cb.d_rt); cudaFree(cb.d_hvec); cudaFree(cb.d_first); cudaFree(cb.d_r0); cudaFree(cb.d_em_scratch);
  cudaFree(cb.d_dense); cudaFree(cb.d_dense_i); cudaFree(cb.d_toff); cudaFree(cb.d_cfgT);
  for (int k = 0; k < 3; k++) {
    if (cb.dense_stream[k]) cudaStreamDestroy(cb.dense_stream[k]);
    if (cb.dense_join[k]) cudaEventDestroy(cb.dense_join[k]);
  }
  if (cb.dense_fork) cudaEventDestroy(cb.dense_fork);
  cb = ChainBatch();
}

// per-call evidence plumbing: static lambda, combined table, per-row configuration.
// Cached on the batch: a repeated call with the same plan only launches the hot kernels.
static int chain_prepare_evidence(const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan,
                                  const ChainInferArgs& a, cudaStream_t st) {
  std::vector<int> key;
  key.insert(key.end(), plan.active_leaf.begin(), plan.active_leaf.end());
  key.push_back(-1);
  key.insert(key.end(), plan.col_leaf_slot.begin(), plan.col_leaf_slot.end());
  const bool same_plan = cb.plan_key == key && cb.d_comb && cb.d_rows;
  if (same_plan && cb.plan_version == cm.param_version) return NIPGPU_OK;
  const int na = plan.n_active;
  if (!same_plan) {
    std::vector<long long> inactive_rows;
    for (int l = 0; l < cm.n_real; l++)
      if (std::find(plan.active_leaf.begin(), plan.active_leaf.end(), l) == plan.active_leaf.end())
        inactive_rows.push_back(cm.leaves[l].lam_off + (long long)cm.leaves[l].miss_cfg * cm.SP);
    cudaFree(cb.d_rows);
    cb.d_rows = nullptr;
    if (int e = upload(&cb.d_rows, inactive_rows, st)) return e;
    NIPGPU_CUDA(cudaStreamSynchronize(st));   // `inactive_rows` dies here
    cb.n_inactive = (int)inactive_rows.size();
  }
  // the evidence tables follow the parameters (every M-step); the per-row configuration does not
  k_chain_lam_prod<<<(cm.SP + 127) / 128, 128, 0, st>>>(cm.d_lam, cb.d_rows, cb.n_inactive, cm.S, cm.SP,
                                                        cb.d_lam_static);
  NIPGPU_LAUNCHED();
  ChainComb K;
  for (int i = 0; i < 8; i++) {
    K.mult[i] = i < na ? plan.mult[i] : 1;
    K.n_cfg[i] = i < na ? cm.leaves[plan.active_leaf[i]].n_cfg : 1;
    K.lam_off[i] = i < na ? cm.leaves[plan.active_leaf[i]].lam_off : 0;
  }
  const size_t need = (size_t)plan.n_comb * cm.SP;
  if (cb.comb_cap < need) {
    cudaFree(cb.d_comb);
    cb.d_comb = nullptr;
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_comb, need * sizeof(double)));
    cb.comb_cap = need;
  }
  k_chain_combine<<<(unsigned)((need + 255) / 256), 256, 0, st>>>(cm.d_lam, cb.d_lam_static, na, K,
                                                                 plan.n_comb, cm.S, cm.SP, cb.d_comb);
  NIPGPU_LAUNCHED();
  if (!same_plan) {
    std::vector<int> cols;  // slot | stride | card | mult
    cols.insert(cols.end(), plan.col_leaf_slot.begin(), plan.col_leaf_slot.end());
    cols.insert(cols.end(), plan.col_stride.begin(), plan.col_stride.end());
    cols.insert(cols.end(), plan.col_card.begin(), plan.col_card.end());
    cols.insert(cols.end(), plan.col_mult.begin(), plan.col_mult.end());
    cudaFree(cb.d_cols);
    cb.d_cols = nullptr;
    if (int e = upload(&cb.d_cols, cols, st)) return e;
    if (a.rows > 0) {
      k_chain_cfg<<<(unsigned)((a.rows + 255) / 256), 256, 0, st>>>(
          a.d_obs, a.rows, a.n_obs, cb.d_cols, cb.d_cols + a.n_obs, cb.d_cols + 2 * a.n_obs,
          cb.d_cols + 3 * a.n_obs, plan.c_miss, cb.d_cfg);
      NIPGPU_LAUNCHED();
    }
    NIPGPU_CUDA(cudaStreamSynchronize(st));  // host vectors above die here
    cb.plan_key = key;
  }
  cb.plan_version = cm.param_version;
  return NIPGPU_OK;
}

// ---- one-tile interfaces on large batches: one thread per sequence (chain_small.cuh) ----
#ifndef NIPGPU_SMALL_MIN_SERIES
#define NIPGPU_SMALL_MIN_SERIES 2048
#endif
// NIPGPU_CHAIN_SMALL=0 keeps the DMMA kernels, =1 forces the thread-per-sequence kernels
static bool small_wanted(int n_series) {
  static const int forced = [] {
    const char* p = getenv("NIPGPU_CHAIN_SMALL");
    return p ? (p[0] == '0' ? 0 : 1) : -1;
  }();
  if (forced >= 0) return forced == 1;
  return n_series >= NIPGPU_SMALL_MIN_SERIES;
}

template <int S>
static int small_launch(const SmallDev& D, int n_lam, size_t smem, const ChainInferArgs& a, double* alphaT,
                        cudaStream_t st) {
  const int grid = (D.n_series + 127) / 128;
  const bool filt = a.forward_only && a.d_post;
  const bool smooth = !a.forward_only && a.d_post;
  constexpr int unit = SmallGeom<S>::V4 ? 4 : 2;
  const int wide = (S % 2 == 0) && a.d_post && (a.post_stride % unit == 0) && (a.post_off % unit == 0) &&
                   ((uintptr_t)a.d_post % (8 * unit) == 0);
  auto fwd = [&](auto kernel) -> int {
    if (int e = set_smem(kernel, smem)) return e;
    kernel<<<grid, 128, smem, st>>>(D, n_lam, smooth ? 1 : 0, alphaT, a.d_post, a.post_stride, a.post_off, wide,
                                    a.d_ll, a.d_status);
    NIPGPU_LAUNCHED();
    return NIPGPU_OK;
  };
  int e;
  if (filt) e = a.want_ll ? fwd(k_chain_small_forward<S, true, true>) : fwd(k_chain_small_forward<S, true, false>);
  else e = a.want_ll ? fwd(k_chain_small_forward<S, false, true>) : fwd(k_chain_small_forward<S, false, false>);
  if (e) return e;
  if (g_chain_mid_event) NIPGPU_CUDA(cudaEventRecord(g_chain_mid_event, st));
  if (smooth) {
    if (int e2 = set_smem(k_chain_small_backward<S>, smem)) return e2;
    k_chain_small_backward<S><<<grid, 128, smem, st>>>(D, n_lam, alphaT, a.d_post, a.post_stride, a.post_off, wide);
    NIPGPU_LAUNCHED();
  }
  return NIPGPU_OK;
}

static int small_infer(const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan, const ChainBatchDev& B,
                       const ChainInferArgs& a, cudaStream_t st) {
  if (cb.cfgT_key != cb.plan_key) {   // time-major copy of the evidence index, once per plan
    k_chain_small_cfgT<<<(a.n_series + 127) / 128, 128, 0, st>>>(cb.d_cfg, cb.d_toff, a.d_row_off, cb.d_order,
                                                                 cb.d_len_sorted, a.n_series, cb.d_cfgT);
    NIPGPU_LAUNCHED();
    cb.cfgT_key = cb.plan_key;
  }
  SmallDev D;
  D.S = cm.S; D.SP = cm.SP; D.c_miss = plan.c_miss; D.n_comb = plan.n_comb; D.m1_0 = cm.m1_0;
  D.A = cm.d_As; D.phi0 = cm.d_phi0; D.R1 = cm.d_R1; D.lam_comb = cb.d_comb;
  D.toff = cb.d_toff; D.cfgT = cb.d_cfgT; D.n_series = a.n_series; D.order = B.order;
  D.len_sorted = B.len_sorted; D.row_off = B.row_off;
  const int n_lam = (size_t)plan.n_comb * cm.S * sizeof(double) <= 16384 ? plan.n_comb : 0;
  const int t_longest = a.n_series > 0 ? std::max(cb.len_sorted[0], 0) : 0;
  D.n_toff = t_longest + 1 <= 1024 ? t_longest + 1 : 0;
  const size_t smem = sizeof(double) * ((size_t)cm.S * cm.S + (size_t)n_lam * cm.S + (size_t)D.n_toff);
  switch (cm.S) {
    case 1: return small_launch<1>(D, n_lam, smem, a, cb.d_alpha, st);
    case 2: return small_launch<2>(D, n_lam, smem, a, cb.d_alpha, st);
    case 3: return small_launch<3>(D, n_lam, smem, a, cb.d_alpha, st);
    case 4: return small_launch<4>(D, n_lam, smem, a, cb.d_alpha, st);
    case 5: return small_launch<5>(D, n_lam, smem, a, cb.d_alpha, st);
    case 6: return small_launch<6>(D, n_lam, smem, a, cb.d_alpha, st);
    case 7: return small_launch<7>(D, n_lam, smem, a, cb.d_alpha, st);
    case 8: return small_launch<8>(D, n_lam, smem, a, cb.d_alpha, st);
  }
  set_error("chain: unsupported interface size");
  return NIPGPU_EUNSUPPORTED;
}

// instrumentation: when set, recorded between the forward and the backward kernel of chain_infer
cudaEvent_t g_chain_mid_event = nullptr;

int chain_infer(const HostModel& hm, const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan,
                const ChainInferArgs& a, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
  (void)hm;
  if (int e = chain_prepare_evidence(cm, cb, plan, a, st)) return e;
  ChainBatchDev B;
  B.n_series = a.n_series; B.order = cb.d_order; B.len_sorted = cb.d_len_sorted;
  const bool ranged = a.p1 >= 0 && !cm.dense && cm.NT > 1;   // a range of the length-sorted sequences
  if (ranged) { B.n_series = a.p1 - a.p0; B.order = cb.d_order + a.p0; B.len_sorted = cb.d_len_sorted + a.p0; }
  B.cfg = cb.d_cfg; B.row_off = a.d_row_off;
  B.fexp = cb.d_fexp; B.zc = cb.d_zc; B.zf = cb.d_zf; B.rn_out = cb.d_rn;
  ChainDev C;
  C.S = cm.S; C.SP = cm.SP; C.c_miss = plan.c_miss;
  C.AS = cm.NT == 1 ? std::max(2, 2 * ((cm.S + 1) / 2)) : cm.SP;
  C.Bf1 = cm.d_Bf1; C.Bb1 = cm.d_Bb1; C.Bb0 = cm.d_Bb0; C.phi0 = cm.d_phi0; C.lam0 = cm.d_lam0;
  C.R1 = cm.d_R1; C.colsum = cm.d_colsum; C.m1_0 = cm.m1_0; C.lam_comb = cb.d_comb;
  if (a.n_series == 0) return NIPGPU_OK;
  if (ev0) NIPGPU_CUDA(cudaEventRecord(ev0, st));
  if (cm.dense) {
    if (int e = dense_infer(cm, cb, plan, a, st)) return e;
    if (ev1) NIPGPU_CUDA(cudaEventRecord(ev1, st));
    return NIPGPU_OK;
  }
  int e = NIPGPU_OK;
  if (cm.NT == 1 && small_wanted(a.n_series)) {
    e = small_infer(cm, cb, plan, B, a, st);
    if (e) return e;
    if (ev1) NIPGPU_CUDA(cudaEventRecord(ev1, st));
    return NIPGPU_OK;
  }
  switch (cm.NT) {
    case 1: e = launch_forward<1>(C, B, a, cb.d_alpha, st); break;
    case 2: e = launch_forward<2>(C, B, a, cb.d_alpha, st); break;
    case 4: e = launch_forward<4>(C, B, a, cb.d_alpha, st); break;
    case 8: e = launch_forward<8>(C, B, a, cb.d_alpha, st); break;
    default: set_error("chain: unsupported interface size"); return NIPGPU_EUNSUPPORTED;
  }
  if (e) return e;
  if (g_chain_mid_event) NIPGPU_CUDA(cudaEventRecord(g_chain_mid_event, st));
  if (!a.forward_only && a.d_post) {
    switch (cm.NT) {
      case 1: e = launch_backward<1>(C, B, a, cb.d_alpha, st); break;
      case 2: e = launch_backward<2>(C, B, a, cb.d_alpha, st); break;
      case 4: e = launch_backward<4>(C, B, a, cb.d_alpha, st); break;
      case 8: e = launch_backward<8>(C, B, a, cb.d_alpha, st); break;
    }
    if (e) return e;
  }
  if (ev1) NIPGPU_CUDA(cudaEventRecord(ev1, st));
  return NIPGPU_OK;
}

// ---- queries beyond the interface variables (see chain.cuh) ----------------------------
namespace {

// W[cfg][ip][y] = sum over the leaf's free-variable combinations r with digit_slot(r) == y that
// are compatible with cfg of leaf_table[base[s(ip)] + off[r]], divided by Lambda[cfg][ip]
__global__ void k_chain_leafpost(const double* leaf_tab, const int* base, const int* off, int R,
                                 const int* ip_to_s, const int* meta, int slot, int card_y, int n_cfg, int S,
                                 int SP, const double* lam, double* W) {
  const long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (x >= (long long)n_cfg * S * card_y) return;
  const int y = (int)(x % card_y), ip = (int)((x / card_y) % S), cfg = (int)(x / ((long long)card_y * S));
  const int nf = meta[0];
  const int* card = meta + 1;
  const int* stride = meta + 1 + nf;
  const int b = base[ip_to_s[ip]];
  double s = 0;
  for (int r = 0; r < R; r++) {
    int rem = r, ok = 1;
    for (int k = 0; k < nf; k++) {
      const int digit = rem % card[k];
      rem /= card[k];
      const int code = (cfg / stride[k]) % (card[k] + 1);
      if (code != card[k] && code != digit) ok = 0;
      if (k == slot && digit != y) ok = 0;
    }
    if (ok) s += leaf_tab[b + off[r]];
  }
  const double l = lam[(long long)cfg * SP + ip];
  W[x] = l != 0 ? s / l : 0.0;
}

// first slices: gprev[series][i] = sum_j gamma_0(j) base0(i, j) / phi0(j)
__global__ void k_chain_prev_first(const double* joint, const long long* row_off, int n_series, long long rows,
                                   const double* base0, const int* ent_of, const double* phi0, int S, int SP,
                                   double* gprev) {
  const long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (x >= (long long)n_series * S) return;
  const int s = (int)(x / S), i = (int)(x - (long long)s * S);
  const long long r = row_off[s];
  double acc = 0;
  if (r < rows) {
    const double* g = joint + r * SP;
    for (int j = 0; j < S; j++) {
      const double p = phi0[j];
      if (p != 0) acc += g[j] * base0[ent_of[i * S + j]] / p;
    }
  }
  gprev[(long long)s * SP + i] = acc;
}

struct PostVarsDev {
  int n;
  int kind[16], stride[16], card[16], off[16], mult[16], n_cfg[16], fixed_cfg[16];
  long long w_off[16];
};

// one thread per (row, output column)
__global__ void k_chain_post_vars(const double* __restrict__ joint, const int* __restrict__ cfg,
                                  const unsigned char* __restrict__ first, long long rows, int S, int SP,
                                  PostVarsDev Q, const double* __restrict__ Wall, int out_row,
                                  double* __restrict__ out) {
  const long long total = rows * out_row;
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const long long r = x / out_row;
    const int col = (int)(x - r * out_row);
    int q = 0;
    while (q + 1 < Q.n && Q.off[q + 1] <= col) q++;
    const int d = col - Q.off[q];
    double s = 0;
    if (Q.kind[q] == 2) {
      const int lc = Q.n_cfg[q] > 0 ? (cfg[r] / Q.mult[q]) % Q.n_cfg[q] : Q.fixed_cfg[q];
      const double* W = Wall + Q.w_off[q] + (long long)lc * S * Q.card[q] + d;
      const double* g = joint + r * SP;
      for (int st = 0; st < S; st++) s += g[st] * W[(long long)st * Q.card[q]];
    } else {
      if (Q.kind[q] == 1 && first[r]) continue;   // written by k_chain_post_first
      const double* g = joint + (Q.kind[q] == 1 ? r - 1 : r) * SP;
      const int stride = Q.stride[q], card = Q.card[q];
      for (int hi = d * stride; hi < S; hi += stride * card)
        for (int lo = 0; lo < stride; lo++) s += g[hi + lo];
    }
    out[x] = s;
  }
}

// previous-slice interface variables on the first slice of every series, from gprev
__global__ void k_chain_post_first(const double* __restrict__ gprev, const long long* __restrict__ row_off,
                                   int n_series, long long rows, int S, int SP, PostVarsDev Q, int out_row,
                                   double* __restrict__ out) {
  const long long total = (long long)n_series * out_row;
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const int sidx = (int)(x / out_row), col = (int)(x - (long long)sidx * out_row);
    const long long r = row_off[sidx];
    if (r >= rows) continue;
    int q = 0;
    while (q + 1 < Q.n && Q.off[q + 1] <= col) q++;
    if (Q.kind[q] != 1) continue;
    const int d = col - Q.off[q], stride = Q.stride[q], card = Q.card[q];
    const double* g = gprev + (long long)sidx * SP;
    double s = 0, tot = 0;
    for (int i = 0; i < S; i++) {
      tot += g[i];
      if ((i / stride) % card == d) s += g[i];
    }
    out[r * out_row + col] = tot != 0 ? s / tot : 0.0;
  }
}

}  // namespace

bool chain_query_plan(const HostModel& hm, const ChainModel& cm, const ChainPlan& plan, int nq,
                      const int32_t* query, int forward_only, std::vector<ChainQueryVar>& out) {
  out.clear();
  if (nq <= 0 || nq > 16 || hm.nif <= 0) return false;
  int off = 0;
  long long w_off = 0;
  for (int i = 0; i < nq; i++) {
    const int v = query[i];
    ChainQueryVar q{};
    q.var = v; q.card = hm.card[v]; q.off = off; q.kind = -1;
    int stride = 1;
    for (int k = 0; k < hm.nif; k++) {
      if (hm.outg[k] == v) { q.kind = 0; q.stride = stride; }
      if (hm.prev[k] == v) { q.kind = 1; q.stride = stride; }
      stride *= hm.card[hm.outg[k]];
    }
    if (q.kind == 1 && forward_only) return false;   // filtering would need a one-step smoother
    if (q.kind < 0) {
      const int l = cm.var_leaf[v];
      if (l < 0 || l >= cm.n_real) return false;
      q.kind = 2; q.leaf = l; q.slot = cm.var_slot[v];
      const int a = (int)(std::find(plan.active_leaf.begin(), plan.active_leaf.end(), l) - plan.active_leaf.begin());
      if (a < (int)plan.active_leaf.size()) { q.mult = plan.mult[a]; q.n_cfg = cm.leaves[l].n_cfg; }
      else { q.mult = 1; q.n_cfg = 0; q.fixed_cfg = cm.leaves[l].miss_cfg; }
      q.w_off = w_off;
      w_off += (long long)cm.leaves[l].n_cfg * cm.S * q.card;
      if (w_off > (1LL << 27)) return false;
    }
    off += q.card;
    out.push_back(q);
  }
  return true;
}

int chain_post_vars(const HostModel& hm, const ChainModel& cm, const ChainBatch& cb,
                    const std::vector<ChainQueryVar>& qv, const double* d_base0, const double* d_base1,
                    const std::vector<int>& tab_off, const int* d_ipool, const double* joint,
                    const unsigned char* first, const long long* d_row_off, int n_series, long long rows,
                    int out_row, double* out, cudaStream_t st) {
  if (rows <= 0 || out_row <= 0) return NIPGPU_OK;
  const int S = cm.S, SP = cm.SP;
  PostVarsDev Q{};
  Q.n = (int)qv.size();
  long long w_total = 0;
  bool any_prev = false;
  for (int i = 0; i < Q.n; i++) {
    Q.kind[i] = qv[i].kind; Q.stride[i] = qv[i].stride; Q.card[i] = qv[i].card; Q.off[i] = qv[i].off;
    Q.mult[i] = std::max(qv[i].mult, 1); Q.n_cfg[i] = qv[i].n_cfg; Q.fixed_cfg[i] = qv[i].fixed_cfg; Q.w_off[i] = qv[i].w_off;
    if (qv[i].kind == 2) w_total = std::max(w_total, qv[i].w_off + (long long)cm.leaves[qv[i].leaf].n_cfg * S * qv[i].card);
    if (qv[i].kind == 1) any_prev = true;
  }
  double* d_W = nullptr;
  double* d_gprev = nullptr;
  std::vector<int*> metas;
  auto cleanup = [&]() { cudaFree(d_W); cudaFree(d_gprev); for (int* p : metas) cudaFree(p); };
  if (w_total > 0 && cudaMalloc((void**)&d_W, (size_t)w_total * sizeof(double)) != cudaSuccess) return NIPGPU_ENOMEM;
  for (int i = 0; i < Q.n; i++) {
    if (qv[i].kind != 2) continue;
    const ChainLeafHost& L = cm.leaves[qv[i].leaf];
    const Proj& p = hm.projs[L.proj];
    std::vector<int> meta{(int)L.free_vars.size()};
    for (int v : L.free_vars) meta.push_back(hm.card[v]);
    for (int sstride : L.cfg_stride) meta.push_back(sstride);
    int* d_meta = nullptr;
    if (cudaMalloc((void**)&d_meta, meta.size() * sizeof(int)) != cudaSuccess) { cleanup(); return NIPGPU_ENOMEM; }
    metas.push_back(d_meta);
    cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);   // `meta` dies at the end of this iteration
    const long long n = (long long)L.n_cfg * S * qv[i].card;
    k_chain_leafpost<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
        d_base1 + tab_off[L.clique], d_ipool + p.base_pos, d_ipool + p.off_pos, p.R,
        cm.d_ip_to_s + (size_t)qv[i].leaf * S, d_meta, qv[i].slot, qv[i].card, L.n_cfg, S, SP,
        cm.d_lam + L.lam_off, d_W + qv[i].w_off);
    NIPGPU_LAUNCHED();
  }
  const long long total = rows * out_row;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
  k_chain_post_vars<<<grid, 256, 0, st>>>(joint, cb.d_cfg, first, rows, S, SP, Q, d_W, out_row, out);
  NIPGPU_LAUNCHED();
  if (any_prev && n_series > 0) {
    if (cudaMalloc((void**)&d_gprev, (size_t)n_series * SP * sizeof(double)) != cudaSuccess) { cleanup(); return NIPGPU_ENOMEM; }
    const long long n = (long long)n_series * S;
    k_chain_prev_first<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(joint, d_row_off, n_series, rows,
                                                                    d_base0 + tab_off[cm.c0], cm.d_ent_of, cm.d_phi0,
                                                                    S, SP, d_gprev);
    NIPGPU_LAUNCHED();
    const long long t2 = (long long)n_series * out_row;
    k_chain_post_first<<<(unsigned)std::min<long long>((t2 + 255) / 256, 148 * 32), 256, 0, st>>>(
        d_gprev, d_row_off, n_series, rows, S, SP, Q, out_row, out);
    NIPGPU_LAUNCHED();
  }
  const cudaError_t err = cudaStreamSynchronize(st);
  cleanup();
  if (err != cudaSuccess) { set_error(std::string("chain_post_vars: ") + cudaGetErrorString(err)); return NIPGPU_ECUDA; }
  return NIPGPU_OK;
}

// ---- ancestral sampling ------------------------------------------------------------------
namespace {

struct SampleLeaf {
  int tab, base, off, R, n_free, ip_to_s;   // offsets into base1 / the int pool / d_ip_to_s
  int card[8], var[8];
};
struct SamplePlan {
  int S, nif, nv, n_leaves, c0_tab;
  int out_var[16], prev_var[16], card[16];
};

__device__ __forceinline__ double uniform01(unsigned long long& x) {   // splitmix64
  x += 0x9e3779b97f4a7c15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// index i with probability w(i) / sum w over n entries given by f(i)
template <class F>
__device__ __forceinline__ int draw(int n, F f, unsigned long long& rng) {
  double tot = 0;
  for (int i = 0; i < n; i++) tot += f(i);
  const double u = uniform01(rng) * tot;
  double acc = 0;
  int last = 0;
  for (int i = 0; i < n; i++) {
    const double w = f(i);
    if (w > 0) last = i;
    acc += w;
    if (u < acc && w > 0) return i;
  }
  return last;
}

__global__ void k_chain_sample(SamplePlan Q, const SampleLeaf* leaves, const double* base0, const double* base1,
                               const int* ipool, const int* ent_of, const int* ip_to_s, const double* phi0,
                               const double* lam0, int n_series, int length, unsigned long long seed, int* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_series) return;
  unsigned long long rng = seed * 0x2545f4914f6cdd1dull + (unsigned long long)s * 0xd1342543de82ef95ull + 1;
  const int S = Q.S;
  const double* c0_0 = base0 + Q.c0_tab;
  const double* c0_1 = base1 + Q.c0_tab;
  // lam0[j] = mass of the slice's leaves without evidence: 1 for proper CPTs, and what the
  // reference's per-slice propagation weighs the state with otherwise (generate_data samples every
  // variable from its marginal in the consistent, partially instantiated slice)
  int state = draw(S, [&](int j) { return phi0[j] * lam0[j]; }, rng);
  int prev = draw(S, [&](int i) { return c0_0[ent_of[i * S + state]]; }, rng);
  for (int t = 0; t < length; t++) {
    if (t > 0) {
      prev = state;
      state = draw(S, [&](int j) { return c0_1[ent_of[prev * S + j]] * lam0[j]; }, rng);
    }
    int* row = out + ((long long)s * length + t) * Q.nv;
    for (int k = 0, st = state, pv = prev; k < Q.nif; k++) {
      row[Q.out_var[k]] = st % Q.card[k];
      row[Q.prev_var[k]] = pv % Q.card[k];
      st /= Q.card[k];
      pv /= Q.card[k];
    }
    for (int l = 0; l < Q.n_leaves; l++) {
      const SampleLeaf L = leaves[l];
      const double* T = base1 + L.tab + ipool[L.base + ip_to_s[L.ip_to_s + state]];
      int r = draw(L.R, [&](int x) { return T[ipool[L.off + x]]; }, rng);
      for (int k = 0; k < L.n_free; k++) {
        row[L.var[k]] = r % L.card[k];
        r /= L.card[k];
      }
    }
  }
}

}  // namespace

int chain_sample(const HostModel& hm, const ChainModel& cm, const double* d_base0, const double* d_base1,
                 const std::vector<int>& tab_off, const int* d_ipool, int n_series, int length,
                 unsigned long long seed, int* d_out, cudaStream_t st) {
  if (!cm.ok || hm.nif > 16) return NIPGPU_EUNSUPPORTED;
  SamplePlan Q{};
  Q.S = cm.S; Q.nif = hm.nif; Q.nv = hm.nv; Q.n_leaves = cm.n_real; Q.c0_tab = tab_off[cm.c0];
  std::vector<char> covered(hm.nv, 0);
  for (int k = 0; k < hm.nif; k++) {
    Q.out_var[k] = hm.outg[k]; Q.prev_var[k] = hm.prev[k]; Q.card[k] = hm.card[hm.outg[k]];
    covered[hm.outg[k]] = covered[hm.prev[k]] = 1;
  }
  std::vector<SampleLeaf> leaves(std::max(cm.n_real, 1));
  for (int l = 0; l < cm.n_real; l++) {
    const ChainLeafHost& L = cm.leaves[l];
    const Proj& p = hm.projs[L.proj];
    if (L.free_vars.size() > 8) return NIPGPU_EUNSUPPORTED;
    SampleLeaf& d = leaves[l];
    d.tab = tab_off[L.clique]; d.base = p.base_pos; d.off = p.off_pos; d.R = p.R;
    d.n_free = (int)L.free_vars.size(); d.ip_to_s = l * cm.S;
    // off[r] enumerates the leaf's free variables in clique order, first fastest (HostModel::add_proj)
    for (int k = 0; k < d.n_free; k++) { d.card[k] = hm.card[L.free_vars[k]]; d.var[k] = L.free_vars[k]; covered[L.free_vars[k]] = 1; }
  }
  for (int v = 0; v < hm.nv; v++)
    if (!covered[v]) return NIPGPU_EUNSUPPORTED;
  if (n_series <= 0 || length <= 0) return NIPGPU_OK;
  SampleLeaf* d_leaves = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d_leaves, leaves.size() * sizeof(SampleLeaf)));
  NIPGPU_CUDA(cudaMemcpyAsync(d_leaves, leaves.data(), leaves.size() * sizeof(SampleLeaf), cudaMemcpyHostToDevice, st));
  k_chain_sample<<<(n_series + 127) / 128, 128, 0, st>>>(Q, d_leaves, d_base0, d_base1, d_ipool, cm.d_ent_of,
                                                         cm.d_ip_to_s, cm.d_phi0, cm.d_lam0, n_series, length, seed,
                                                         d_out);
  g_launches++;
  const cudaError_t err = cudaStreamSynchronize(st);
  cudaFree(d_leaves);
  if (err != cudaSuccess) { set_error(std::string("chain_sample: ") + cudaGetErrorString(err)); return NIPGPU_ECUDA; }
  return NIPGPU_OK;
}

// E-step of a whole batch on the chain engine: forward, backward (EM flavour), transition
// counts as a DMMA GEMM, leaf counts, then expected clique tables -> per-variable family
// counts in the layout of em_learn's `parameters[]` (src/nip.c:2108-2128).
// Returns NIPGPU_EUNSUPPORTED when the evidence layout does not fit (caller uses engine 1).
int chain_estep(const HostModel& hm, const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan,
                const ChainEmArgs& x, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
  const ChainInferArgs& a = x.base;
  const int SP = cm.SP, S = cm.S;
  const bool dense = cm.dense;   // |I| > 64: per-slice GEMMs (dense.cu) + dense_stats
  const size_t tab = (size_t)plan.n_comb * SP;
  // k_chain_stats: two (own, beta) tile stages + posterior tile + `phases` evidence-indexed tables
  const size_t stats_fixed = sizeof(double) * (6 * 33 * (SP + 4) + 2 * (3 * 34 + 1) + tab) + 3 * 34 * sizeof(int) + 16;
  int phases = 4;
  while (phases > 1 && stats_fixed + phases * tab * sizeof(double) > 220 * 1024) phases /= 2;
  if (!dense && stats_fixed + phases * tab * sizeof(double) > 220 * 1024) return NIPGPU_EUNSUPPORTED;
  if (dense && (size_t)plan.n_comb * 128 * sizeof(double) > 200 * 1024) return NIPGPU_EUNSUPPORTED;
  for (int l = 0; l < cm.n_real; l++)
    if (cm.leaves[l].free_vars.size() > 8) return NIPGPU_EUNSUPPORTED;
  if (int e = chain_prepare_evidence(cm, cb, plan, a, st)) return e;

  const long long rows = std::max<long long>(a.rows, 1);
  // CTAs of the statistics kernel (one per SM, split-K); dense: split-K of the count GEMM, the
  // leaf pass uses row ranges so that its grid fills the machine
  int parts = std::max(1, x.sm_count);
  if (dense) {  // split-K of the count GEMM: fill whole waves of SMs (1024 states: 64 tiles x 37 = 16 x 148)
    const int tiles = (SP / 128) * (SP / 128), sms = std::max(1, x.sm_count);
    double best = 0;
    parts = 8;
    for (int k = 4; k <= 48; k++) {
      const int ctas = tiles * k, waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / ((double)waves * sms);
      if (eff > best + 1e-9) { best = eff; parts = k; }
    }
  }
  const int partsC = dense ? std::max(1, 2 * x.sm_count / (SP / 128)) : parts;
  if (!cb.d_rt) {
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_rt, rows * SP * sizeof(double)));
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_hvec, rows * sizeof(double)));
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_rn, rows * sizeof(double)));
    NIPGPU_CUDA(cudaMemsetAsync(cb.d_rn, 0, rows * sizeof(double), st));
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_first, rows));
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_r0, (size_t)std::max(a.n_series, 1) * SP * sizeof(double)));
    NIPGPU_CUDA(cudaMemsetAsync(cb.d_first, 0, rows, st));
    NIPGPU_CUDA(cudaMemsetAsync(cb.d_hvec, 0, rows * sizeof(double), st));
    if (a.n_series > 0) {
      k_chain_first<<<(a.n_series + 255) / 256, 256, 0, st>>>(a.d_row_off, a.n_series, a.rows, cb.d_first);
      NIPGPU_LAUNCHED();
    }
  }
  // scratch: part_G | G | g0 | part_C | Cc | E | E0 | work | part_g0
  const size_t n_partG = (size_t)parts * SP * SP, n_G = (size_t)SP * SP, n_partC = (size_t)partsC * tab;
  const size_t n_E = (size_t)hm.toff[hm.nc], n_E0 = (size_t)hm.csize[cm.c0];
  const size_t n_work = dense ? 2 * (size_t)rows : 0;   // dense_stats: N_k and pair weights
  const int g0_parts = std::max(1, std::min(64, (a.n_series + 63) / 64));
  const size_t need = n_partG + n_G + SP + n_partC + tab + n_E + n_E0 + n_work + (size_t)g0_parts * SP;
  if (cb.em_scratch_cap < need) {
    cudaFree(cb.d_em_scratch);
    cb.d_em_scratch = nullptr;
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_em_scratch, need * sizeof(double)));
    cb.em_scratch_cap = need;
  }
  double* partG = cb.d_em_scratch;
  double* G = partG + n_partG;
  double* g0 = G + n_G;
  double* partC = g0 + SP;
  double* Cc = partC + n_partC;
  double* E = Cc + tab;
  double* E0 = E + n_E;
  double* work = E0 + n_E0;
  double* part_g0 = work + n_work;

  ChainBatchDev B;
  B.n_series = a.n_series; B.order = cb.d_order; B.len_sorted = cb.d_len_sorted;
  B.cfg = cb.d_cfg; B.row_off = a.d_row_off;
  B.fexp = cb.d_fexp; B.zc = cb.d_zc; B.zf = cb.d_zf; B.rn_out = cb.d_rn;
  ChainDev C;
  C.S = S; C.SP = SP; C.c_miss = plan.c_miss;
  C.AS = cm.NT == 1 ? std::max(2, 2 * ((S + 1) / 2)) : SP;
  C.Bf1 = cm.d_Bf1; C.Bb1 = cm.d_Bb1; C.Bb0 = cm.d_Bb0; C.phi0 = cm.d_phi0; C.lam0 = cm.d_lam0;
  C.R1 = cm.d_R1; C.colsum = cm.d_colsum; C.m1_0 = cm.m1_0; C.lam_comb = cb.d_comb;
  NIPGPU_CUDA(cudaMemsetAsync(cb.d_r0, 0, (size_t)std::max(a.n_series, 1) * SP * sizeof(double), st));
  if (ev0) NIPGPU_CUDA(cudaEventRecord(ev0, st));
  int e = NIPGPU_OK;
  ChainStatsArgs sa;
  sa.bt = cb.d_rt; sa.hv = cb.d_hvec; sa.r0 = cb.d_r0; sa.first = cb.d_first; sa.rows = a.rows;
  sa.n_comb = plan.n_comb; sa.phases = phases; sa.parts = parts;
  sa.smem = stats_fixed + phases * tab * sizeof(double);
  sa.partG = partG; sa.partC = partC;
  if (a.n_series > 0 && dense) {
    DenseEm em;
    em.bt = cb.d_rt; em.hvec = cb.d_hvec; em.r0 = cb.d_r0;
    ChainInferArgs fa = a;
    fa.want_ll = 1; fa.forward_only = 0; fa.d_post = nullptr;
    if ((e = dense_infer(cm, cb, plan, fa, st, &em)) ||
        (e = dense_stats(cm, cb, plan, em, cb.d_first, a.rows, work, parts, partG, partsC, partC, st)))
      return e;
  } else if (a.n_series > 0) {
    switch (cm.NT) {
      case 1: e = launch_em<1>(C, B, a, cb.d_alpha, sa, st); break;
      case 2: e = launch_em<2>(C, B, a, cb.d_alpha, sa, st); break;
      case 4: e = launch_em<4>(C, B, a, cb.d_alpha, sa, st); break;
      case 8: e = launch_em<8>(C, B, a, cb.d_alpha, sa, st); break;
      default: return NIPGPU_EUNSUPPORTED;
    }
    if (e) return e;
  } else {
    NIPGPU_CUDA(cudaMemsetAsync(partG, 0, (n_partG + n_G + SP + n_partC) * sizeof(double), st));
  }
  k_chain_sum_parts<<<(unsigned)((n_G + 31) / 32), 256, 0, st>>>(partG, parts, (long long)n_G, G);
  NIPGPU_LAUNCHED();
  k_chain_g0_part<<<g0_parts, 256, 0, st>>>(cb.d_r0, a.n_series, SP, (a.n_series + g0_parts - 1) / g0_parts, part_g0);
  NIPGPU_LAUNCHED();
  k_chain_sum_parts<<<(unsigned)((SP + 31) / 32), 256, 0, st>>>(part_g0, g0_parts, (long long)SP, g0);
  NIPGPU_LAUNCHED();
  k_chain_sum_parts<<<(unsigned)((tab + 31) / 32), 256, 0, st>>>(partC, partsC, (long long)tab, Cc);
  NIPGPU_LAUNCHED();
  if (ev1) NIPGPU_CUDA(cudaEventRecord(ev1, st));
  // ---- expected clique tables ----
  const int n0 = hm.csize[cm.c0];
  k_chain_expect_c0<<<(n0 + 255) / 256, 256, 0, st>>>(x.d_base0 + (*x.tab_off)[cm.c0], x.d_base1 + (*x.tab_off)[cm.c0],
                                                      cm.d_ent_im, cm.d_ent_ip, n0, SP, G, g0,
                                                      E + (*x.tab_off)[cm.c0], E0);
  NIPGPU_LAUNCHED();
  for (int l = 0; l < cm.n_real; l++) {
    const ChainLeafHost& Lh = cm.leaves[l];
    const Proj& p = hm.projs[Lh.proj];
    LeafExpect L;
    L.n_free = (int)Lh.free_vars.size();
    for (int k = 0; k < L.n_free; k++) { L.card[k] = hm.card[Lh.free_vars[k]]; L.stride[k] = Lh.cfg_stride[k]; }
    L.slot = (int)(std::find(plan.active_leaf.begin(), plan.active_leaf.end(), l) - plan.active_leaf.begin());
    if (L.slot == (int)plan.active_leaf.size()) L.slot = -1;
    L.mult = L.slot >= 0 ? plan.mult[L.slot] : 1;
    L.n_cfg = Lh.n_cfg; L.miss_cfg = Lh.miss_cfg;
    L.m = p.m; L.R = p.R; L.S = S; L.SP = SP; L.n_comb = plan.n_comb; L.lam_off = Lh.lam_off;
    const int n = p.m * p.R;
    k_chain_expect_leaf<<<(n + 3) / 4, 128, 0, st>>>(x.d_base1 + (*x.tab_off)[Lh.clique], x.d_ipool + p.base_pos,
                                                         x.d_ipool + p.off_pos, cm.d_ip_to_s + (size_t)l * S, L, Cc,
                                                         cm.d_lam, E + (*x.tab_off)[Lh.clique]);
    NIPGPU_LAUNCHED();
  }
  // ---- family counts of every variable ----
  for (int v = 0; v < hm.nv; v++) {
    const Proj& p = hm.projs[hm.proj_fam[v]];
    const int c = hm.family[v];
    const bool first_only = (hm.flags[v] & NIPGPU_IF_OLD_OUTGOING) != 0;   // src/nip.c:1932
    const double* src = (first_only && c == cm.c0) ? E0 : E + (*x.tab_off)[c];
    k_chain_family<<<(p.m + 127) / 128, 128, 0, st>>>(src, x.d_ipool + p.base_pos, x.d_ipool + p.off_pos, p.m,
                                                      p.R, x.pseudo, x.d_counts + hm.coff[v]);
    NIPGPU_LAUNCHED();
  }
  k_chain_tail<<<1, 512, 0, st>>>(a.d_ll, a.d_status, a.n_series, x.d_counts + hm.coff[hm.nv]);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

}  // namespace nipgpu
