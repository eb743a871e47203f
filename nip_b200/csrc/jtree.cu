// jtree.cu — generic join-tree engine: hand-written sm_100a kernels for the
// potential algebra (marginalise / multiply-divide update / evidence /
// normalise) driven by a per-slice device schedule.
//
// Reference semantics implemented here (file:line in manuelschmidt/nip):
//   op_marg            nip_general_marginalise / nip_total_marginalise  src/nippotential.c:267-346
//   op_absorb(_ratio)  nip_update_potential (0/0 -> 0)                  src/nippotential.c:436-496
//   op_evidence        nip_update_evidence with an indicator vector     src/nippotential.c:499-522,
//                      nip_enter_index_observation                      src/nipjointree.c:832-856
//   vec_normalise      nip_normalise_array (zero sum: untouched)        src/nippotential.c:349-360
//   do_collect/do_distribute  nip_collect_evidence / nip_distribute_evidence
//                      + nip_message_pass                               src/nipjointree.c:580-709
//   k_jt_forward       forward phase of forward(_backward)_inference / e_step
//                      src/nip.c:1435-1493, 1238-1311, 1791-1880
//   k_jt_backward      backward phase                                    src/nip.c:1498-1573, 1885-1990
//   k_jt_likelihood    util/niplikelihood.c:111-135
// The slice always restarts from the shared base tables (original_p x priors),
// which is what reset_model + use_priors produce (src/nip.c:61-119), so
// nip_retract_potential / nip_global_retraction never touch HBM.
//
// Every kernel is written once against a "team" — the set of threads that owns
// one sequence at a time — and instantiated three ways:
//   WarpTeam  one warp per sequence, tables in that warp's slice of shared
//             memory (models whose cliques are a few dozen entries: C1/C5);
//   CtaTeam   one CTA per sequence, tables in shared memory or, when they do
//             not fit, in a per-CTA HBM workspace;
//   GridTeam  the whole (cooperative) grid streams ONE sequence's tables
//             through HBM — cliques of millions of entries (C3: 16^6).
#include "jtree.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>

namespace cg = cooperative_groups;

namespace nipgpu {

namespace {

struct Work {
  double *tab, *msg, *tmp, *va, *scr, *quo;
  // Lazy start of a slice (grid team): a clique nobody has written yet is read straight from
  // the shared base tables, so the 2 x |tables| copy that would open every slice never happens.
  const double* base;
  unsigned long long dirty;  // bit c: clique c has been written and lives in `tab`
  bool lazy;
  __device__ const double* rd_clique(int c) const { return (!lazy || ((dirty >> c) & 1)) ? tab : base; }
  __device__ const double* rd(const DProgram& P, int pj) const { return rd_clique(P.projs[pj].clq); }
  __device__ void wrote(const DProgram& P, int pj) { dirty |= 1ull << P.projs[pj].clq; }
};

__device__ __forceinline__ Work carve(const DProgram& P, double* W) {
  Work w;
  w.tab = W;
  w.msg = w.tab + P.tab_total;
  w.tmp = w.msg + P.msg_total;
  w.va = w.tmp + P.msg_max;
  w.scr = w.va + 3 * P.S;
  w.quo = w.scr + P.scratch;
  w.base = nullptr;
  w.dirty = 0;
  w.lazy = false;
  return w;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kRingSlots = 16;  // 16-byte slots per thread of the grid team's cp.async ring (power of two)
constexpr unsigned long long kTraceCap = (JT_TRACE_WORDS - 1) / 2;  // barriers recorded by NIPGPU_JT_TRACE

struct WarpTeam {
  static constexpr bool kGrid = false;
  static constexpr int kMaxThreads = 256, kMinBlocks = 2;
  static constexpr int kUnroll = 2;
  __device__ WarpTeam(double*, double*, double*, double*) {}
  __device__ int tid() const { return threadIdx.x & 31; }
  __device__ int size() const { return 32; }
  __device__ void mark(unsigned, int) {}
  __device__ void sync() const { __syncwarp(); }
  __device__ double sum(double v) { return warp_sum(v); }
  __device__ int slot() const { return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); }
  __device__ int first() const { return slot(); }
  __device__ int step() const { return gridDim.x * (blockDim.x >> 5); }
  __device__ double* work(double* gwork, double* smem, size_t wstride) const {
    return smem + (size_t)(threadIdx.x >> 5) * wstride;
  }
};

struct CtaTeam {
  static constexpr bool kGrid = false;
  static constexpr int kMaxThreads = 512, kMinBlocks = 1;   // one big CTA when the tables fill an SM's shared memory
  static constexpr int kUnroll = 4;
  double* red;
  __device__ CtaTeam(double* r, double*, double*, double*) : red(r) {}
  __device__ int tid() const { return threadIdx.x; }
  __device__ int size() const { return blockDim.x; }
  __device__ void mark(unsigned, int) {}
  __device__ void sync() const { __syncthreads(); }
  __device__ double sum(double v) { return block_sum(v, red); }
  __device__ int slot() const { return blockIdx.x; }
  __device__ int first() const { return blockIdx.x; }
  __device__ int step() const { return gridDim.x; }
  __device__ double* work(double* gwork, double* smem, size_t wstride) const {
    return gwork ? gwork + (size_t)blockIdx.x * wstride : smem;
  }
};

// All CTAs of a cooperative launch.  `part` is 2 x gridDim doubles (alternating, so one
// grid barrier per sum is enough), `scratch` holds size() doubles for two-stage marginals.
struct GridTeam {
  static constexpr bool kGrid = true;
  static constexpr int kMaxThreads = 256, kMinBlocks = 2;
  static constexpr int kUnroll = 8;
  double *red, *part, *scratch;
  double2* ring;  // this thread's cp.async slots: ring[u * blockDim.x], u < kRingSlots
  int flip = 0;
  // NIPGPU_JT_TRACE=1: trace[0] = records, then (tag, globaltimer ns) per grid barrier
  unsigned long long* trace = nullptr;
  unsigned tag = 0;
  __device__ GridTeam(double* r, double* p, double* s, double* smem)
      : red(r), part(p), scratch(s), ring(reinterpret_cast<double2*>(smem) + threadIdx.x) {}
  __device__ int tid() const { return blockIdx.x * blockDim.x + threadIdx.x; }
  __device__ int size() const { return gridDim.x * blockDim.x; }
  __device__ void mark(unsigned code, int pj) { tag = code << 16 | (unsigned)(pj & 0xffff); }
  __device__ void sync() const {
    cg::this_grid().sync();
    if (trace && tid() == 0) {
      const unsigned long long k = trace[0];
      if (k < kTraceCap) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        trace[1 + 2 * k] = tag;
        trace[2 + 2 * k] = now;
        trace[0] = k + 1;
      }
    }
  }
  __device__ double sum(double v) {
    const double b = block_sum(v, red);
    double* mine = part + flip * gridDim.x;
    flip ^= 1;
    if (threadIdx.x == 0) mine[blockIdx.x] = b;
    sync();
    double s = 0;  // every CTA adds the partials in the same order
    if (threadIdx.x < 32) {
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += mine[i];
      s = warp_sum(s);
      if (threadIdx.x == 0) red[33] = s;
    }
    __syncthreads();
    return red[33];
  }
  int seq_first = 0, seq_step = 1;   // this kernel's share of the sequences (concurrent groups)
  __device__ int slot() const { return seq_first; }
  __device__ int first() const { return seq_first; }
  __device__ int step() const { return seq_step; }
  __device__ double* work(double* gwork, double*, size_t) const { return gwork; }
};

// Start a slice from the shared base tables: a copy into the team's work area, or (grid team,
// at most 64 cliques) nothing at all — see Work::rd.
template <class Team>
__device__ __forceinline__ void load_tables(Team& tm, const DProgram& P, Work& w, const double* __restrict__ src,
                                            bool allow_lazy = true) {
  if (Team::kGrid && allow_lazy && P.n_cliques <= 64) {
    w.base = src;
    w.dirty = 0;
    w.lazy = true;
    return;
  }
  w.lazy = false;
  double* __restrict__ tab = w.tab;
  constexpr int U = Team::kUnroll;
  const int n = P.tab_total, stride = tm.size();
  int i = tm.tid();
  for (; i + (U - 1) * stride < n; i += U * stride) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; u++) tab[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) tab[i] = src[i];
  tm.sync();
}

// Rows of at most PL loads per lane: U destinations at a time, U * PL independent loads in
// flight per thread, the next batch's base offsets fetched while they land.
template <int U, int PL>
__device__ __forceinline__ void marg_rows(GridTeam& tm, const DProj& p, const double* T,
                                          const int* __restrict__ base, const int* __restrict__ off, double* dst) {
  const int lanes = p.lanes, sub = tm.tid() % lanes, grp = tm.tid() / lanes;
  const int groups = tm.size() / lanes;
  int o[PL];
#pragma unroll
  for (int k = 0; k < PL; k++) o[k] = sub + k * lanes < p.R ? off[sub + k * lanes] : -1;
  int b[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int j = u * groups + grp;
    b[u] = j < p.m ? base[j] : -1;
  }
  for (int g0 = 0; g0 < p.m; g0 += groups * U) {
    double v[U][PL];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int k = 0; k < PL; k++) v[u][k] = (b[u] >= 0 && o[k] >= 0) ? T[b[u] + o[k]] : 0.0;
    int bn[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int j = g0 + groups * U + u * groups + grp;
      bn[u] = j < p.m ? base[j] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      double s = v[u][0];
#pragma unroll
      for (int k = 1; k < PL; k++) s += v[u][k];
      for (int w = lanes >> 1; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
      const int j = g0 + u * groups + grp;
      if (sub == 0 && j < p.m) dst[j] = s;
      b[u] = bn[u];
    }
  }
}

// Grid team marginalisation.  The grid has far more threads than a small destination has
// entries: the free range is cut into C chunks per destination entry (partials in scratch,
// chunk-major) and the chunks of each entry are added in a fixed order afterwards.  Every
// thread keeps 16 independent loads in flight, over several destinations when rows are short
// (8-byte cp.async gathers through the ring measured SLOWER than register loads here).
__device__ __noinline__ void grid_marg(GridTeam& tm, const DProgram& P, const double* tab, int pj, double* dst) {
  const DProj p = P.projs[pj];
  const double* T = tab + p.tab;  // written earlier in this kernel: no __restrict__ / ld.global.nc
  const int* __restrict__ base = P.ipool + p.base;
  const int* __restrict__ off = P.ipool + p.off;
  tm.mark(1, pj);
  const int lanes = p.lanes, sub = tm.tid() % lanes, grp = tm.tid() / lanes;
  const int groups = tm.size() / lanes;
  const int per_lane_min = 8;
  const int C = max(1, min(groups / max(p.m, 1), p.R / (lanes * per_lane_min)));
  const int Rc = ((p.R + C - 1) / C + lanes - 1) / lanes * lanes;
  const int total = p.m * C;
  if (C == 1 && Rc <= 8 * lanes) {  // short rows: several destinations per thread at a time
    if (Rc <= lanes) marg_rows<16, 1>(tm, p, T, base, off, dst);
    else if (Rc <= 2 * lanes) marg_rows<8, 2>(tm, p, T, base, off, dst);
    else if (Rc <= 4 * lanes) marg_rows<4, 4>(tm, p, T, base, off, dst);
    else marg_rows<2, 8>(tm, p, T, base, off, dst);
    tm.sync();
    return;
  }
  double* out = C == 1 ? dst : tm.scratch;
  for (int g0 = 0; g0 < total; g0 += groups) {
    const int g = g0 + grp;
    double s = 0;
    int c = 0, j = 0;
    if (g < total) {
      c = g / p.m;
      j = g - c * p.m;
      const double* Tb = T + base[j];
      const int r1 = min(p.R, (c + 1) * Rc);
#pragma unroll 16
      for (int r = c * Rc + sub; r < r1; r += lanes) s += Tb[off[r]];
    }
    for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sub == 0 && g < total) out[c * p.m + j] = s;
  }
  if (C > 1) {
    tm.sync();
    tm.mark(2, pj);
    if (C > 64) {  // a CTA per destination entry
      for (int j = blockIdx.x; j < p.m; j += gridDim.x) {
        double s = 0;
#pragma unroll 4
        for (int c = threadIdx.x; c < C; c += blockDim.x) s += tm.scratch[c * p.m + j];
        s = block_sum(s, tm.red);
        if (threadIdx.x == 0) dst[j] = s;
      }
    } else {
      const int lane = tm.tid() & 31, w = tm.tid() >> 5, nw = tm.size() >> 5;
      for (int j = w; j < p.m; j += nw) {
        double s = 0;
        for (int c = lane; c < C; c += 32) s += tm.scratch[c * p.m + j];
        s = warp_sum(s);
        if (lane == 0) dst[j] = s;
      }
    }
  }
  tm.sync();
}

// Grid team elementwise pass in table order: D[e] = S[e] * cvec[j(e)], or (cvec == nullptr)
// D[e] = j(e) == state ? S[e] : 0.  Every thread keeps kRingSlots 16-byte cp.async copies of
// table entries in flight — global -> its own shared-memory slots, no registers held — which
// is what covers HBM latency at 6+ TB/s.  j(e) comes from the projection's table-order map; the
// coefficients of four slots are gathered together (plain loads, L1/L2 hits) while the copies
// land.
__device__ __noinline__ void grid_update(GridTeam& tm, const DProgram& P, double* dst_tab, const double* src_tab,
                                         int pj, const double* cvec, int state) {
  tm.mark(3, pj);
  const DProj p = P.projs[pj];
  const int n = p.m * p.R;
  double* D = dst_tab + p.tab;
  const double* S = src_tab + p.tab;  // may be D
  const int* __restrict__ jlo = P.ipool + p.jlo;
  const int* __restrict__ jhi = P.ipool + p.jhi;
  const int Fd = p.F, fshift = (Fd & (Fd - 1)) == 0 ? __ffs(Fd) - 1 : -1;
  const int stride = tm.size();
  const bool vec = (n & 1) == 0 && (Fd & 1) == 0 && (((size_t)D | (size_t)S) & 15) == 0;
  if (vec) {
    constexpr int K = kRingSlots, B = 4;
    const int n2 = n >> 1;
    const double2* S2 = reinterpret_cast<const double2*>(S);
    double2* D2 = reinterpret_cast<double2*>(D);
    const int ls = blockDim.x;
    double2* vr = tm.ring;
    const int i = tm.tid();
#pragma unroll
    for (int u = 0; u < K; u++) {
      const long long idx = i + (long long)u * stride;
      if (idx < n2) cp_async16(vr + u * ls, S2 + idx);
      cp_async_commit();
    }
    int slot = 0;
    for (long long idx0 = i; idx0 < n2; idx0 += (long long)B * stride) {
      double c0[B], c1[B];
#pragma unroll
      for (int u = 0; u < B; u++) {
        const long long idx = idx0 + (long long)u * stride;
        c0[u] = c1[u] = 0.0;
        if (idx < n2) {  // F is even: e and e + 1 share the high part
          const int e = (int)(idx * 2);
          const int h = fshift >= 0 ? e >> fshift : e / Fd;
          const int l = e - h * Fd, jh = jhi[h];
          const int j0 = jh + jlo[l], j1 = jh + jlo[l + 1];
          if (cvec) {
            c0[u] = cvec[j0];
            c1[u] = cvec[j1];
          } else {
            c0[u] = j0 == state ? 1.0 : 0.0;
            c1[u] = j1 == state ? 1.0 : 0.0;
          }
        }
      }
      cp_async_wait<K - B>();
#pragma unroll
      for (int u = 0; u < B; u++) {
        const long long idx = idx0 + (long long)u * stride;
        const int sl = (slot + u) & (K - 1);
        if (idx < n2) {
          double2 v = vr[sl * ls];
          if (cvec) {
            v.x *= c0[u];
            v.y *= c1[u];
          } else {  // evidence keeps the entry bit for bit (no multiply: NaN/inf stay out)
            v.x = c0[u] != 0.0 ? v.x : 0.0;
            v.y = c1[u] != 0.0 ? v.y : 0.0;
          }
          D2[idx] = v;
        }
        const long long nxt = idx + (long long)K * stride;
        if (nxt < n2) cp_async16(vr + sl * ls, S2 + nxt);
        cp_async_commit();
      }
      slot = (slot + B) & (K - 1);
    }
    cp_async_wait<0>();
  } else {
    for (int e = tm.tid(); e < n; e += stride) {
      const int h = e / Fd;
      const int j = jhi[h] + jlo[e - h * Fd];
      D[e] = cvec ? S[e] * cvec[j] : (j == state ? S[e] : 0.0);
    }
  }
  tm.sync();
}

// dst[j] = sum_r T[base[j] + off[r]] — `lanes` threads share one destination
// entry and combine with a fixed shuffle tree, so the result is deterministic.
template <class Team>
__device__ void op_marg(Team& tm, const DProgram& P, const double* tab, int pj, double* dst) {
  const DProj p = P.projs[pj];
  const double* T = tab + p.tab;  // no __restrict__: written earlier in this kernel (no ld.global.nc)
  const int* __restrict__ base = P.ipool + p.base;
  const int* __restrict__ off = P.ipool + p.off;
  tm.mark(1, pj);
  if constexpr (Team::kGrid) {
    grid_marg(tm, P, tab, pj, dst);
    return;
  } else {
    // a destination smaller than the team (the message to a small clique, a variable's
    // marginal) would leave most threads idle: up to a warp shares each entry then
    int lanes = p.lanes;
    while (lanes < 32 && p.m * lanes * 2 <= tm.size() && lanes * 2 <= p.R) lanes *= 2;
    const int sub = tm.tid() % lanes, grp = tm.tid() / lanes;
    const int groups = tm.size() / lanes;
    for (int j0 = 0; j0 < p.m; j0 += groups) {
      const int j = j0 + grp;
      double s = 0;
      if (j < p.m) {
        const int b = base[j];
        for (int r = sub; r < p.R; r += lanes) s += T[b + off[r]];
      }
      for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (sub == 0 && j < p.m) dst[j] = s;
    }
  }
  tm.sync();
}

// entry x of the projection's (j, r) enumeration, fastest index chosen so that
// consecutive threads touch consecutive addresses; shifts when the divisor is a power of two
struct Decoder {
  int div, shift;
  bool jfast;
  __device__ explicit Decoder(const DProj& p) {
    jfast = p.lanes == 1;
    div = jfast ? p.m : p.R;
    shift = (div & (div - 1)) == 0 ? __ffs(div) - 1 : -1;
  }
  __device__ __forceinline__ void operator()(int x, int& j, int& r) const {
    const int q = shift >= 0 ? x >> shift : x / div;
    const int rem = x - q * div;
    if (jfast) { r = q; j = rem; } else { j = q; r = rem; }
  }
};

// D[e] = f(S[e], coef(j)) for every entry e = base[j] + off[r] of a projection (S may be D).
// U entries per thread are loaded before any is stored, so that a thread keeps U independent
// requests in flight when the tables live in an HBM workspace.
template <int U, class Team, class Coef, class F>
__device__ __forceinline__ void for_entries(Team& tm, const DProgram& P, double* dst_tab,
                                            const double* src_tab, int pj, Coef coef, F f) {
  {
    const DProj p = P.projs[pj];
    double* D = dst_tab + p.tab;
    const double* S = src_tab + p.tab;  // may be D
    const int* __restrict__ base = P.ipool + p.base;
    const int* __restrict__ off = P.ipool + p.off;
    const Decoder dec(p);
    const int n = p.m * p.R, stride = tm.size();
    int x = tm.tid();
    for (; x + (U - 1) * stride < n; x += U * stride) {
      int a[U];
      double v[U];
      decltype(coef(0)) c[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        int j, r;
        dec(x + u * stride, j, r);
        a[u] = base[j] + off[r];
        c[u] = coef(j);
      }
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = S[a[u]];
#pragma unroll
      for (int u = 0; u < U; u++) D[a[u]] = f(v[u], c[u]);
    }
    for (; x < n; x += stride) {
      int j, r;
      dec(x, j, r);
      const int a = base[j] + off[r];
      D[a] = f(S[a], coef(j));
    }
    tm.sync();
  }
}

// T[base[j] + off[r]] *= v[j]
template <class Team>
__device__ void op_absorb(Team& tm, const DProgram& P, Work& w, int pj, const double* v) {
  if constexpr (Team::kGrid) grid_update(tm, P, w.tab, w.rd(P, pj), pj, v, 0);
  else
    for_entries<Team::kUnroll>(tm, P, w.tab, w.rd(P, pj), pj, [=](int j) { return v[j]; },
                             [](double t, double c) { return t * c; });
  w.wrote(P, pj);
}

// T[..] = T[..] * num[j] / den[j], and 0 where den[j] == 0.  The grid team forms the quotient
// vector once (num/den per destination entry instead of a division per table entry); the
// difference to (t * num) / den is one rounding, far inside the 1e-9 gate.  `keep_num_in`, if
// given, receives a copy of num (the distribute pass stores the new message that way).
struct Ratio { double num, den; };
template <class Team>
__device__ void op_absorb_ratio(Team& tm, const DProgram& P, Work& w, int pj, const double* num,
                                const double* den) {
  if constexpr (Team::kGrid) {
    tm.mark(4, pj);
    const int m = P.projs[pj].m;
    for (int j = tm.tid(); j < m; j += tm.size()) {
      const double d = den[j];
      w.quo[j] = d != 0 ? num[j] / d : 0.0;
    }
    tm.sync();
    op_absorb(tm, P, w, pj, w.quo);
  } else {
    for_entries<Team::kUnroll>(tm, P, w.tab, w.rd(P, pj), pj, [=](int j) { return Ratio{num[j], den[j]}; },
                               [](double t, Ratio c) { return c.den != 0 ? (t * c.num) / c.den : 0.0; });
    w.wrote(P, pj);
  }
}

// hard observation: keep only the entries whose state of the variable is `state`
template <class Team>
__device__ void op_evidence(Team& tm, const DProgram& P, Work& w, int pj, int state) {
  if constexpr (Team::kGrid) grid_update(tm, P, w.tab, w.rd(P, pj), pj, nullptr, state);
  else
    for_entries<Team::kUnroll>(tm, P, w.tab, w.rd(P, pj), pj, [=](int j) { return j == state; },
                             [](double t, bool keep) { return keep ? t : 0.0; });
  w.wrote(P, pj);
}

template <class Team>
__device__ double vec_sum(Team& tm, const double* v, int n) {
  tm.mark(5, 0);
  double s = 0;
  for (int i = tm.tid(); i < n; i += tm.size()) s += v[i];
  return tm.sum(s);
}

template <class Team>
__device__ void vec_normalise(Team& tm, double* v, int n) {
  const double s = vec_sum(tm, v, n);
  tm.mark(6, 0);
  if (s != 0)
    for (int i = tm.tid(); i < n; i += tm.size()) v[i] /= s;
  tm.sync();
}

// collect: child -> parent messages in post-order; sepsets start at 1 so the
// absorbed ratio is the message itself.  Messages are kept for distribute.
template <class Team>
__device__ void do_collect(Team& tm, const DProgram& P, Work& w) {
  for (int i = 0; i < P.n_collect; i++) {
    const DMsg m = P.collect[i];
    op_marg(tm, P, w.rd(P, m.proj_src), m.proj_src, w.msg + m.slot);
    op_absorb(tm, P, w, m.proj_dst, w.msg + m.slot);
  }
}

// distribute: parent -> child; the child absorbs new/old where old is the
// message it sent up during collect.
template <class Team>
__device__ void do_distribute(Team& tm, const DProgram& P, Work& w, const DMsg* list, int n) {
  for (int i = 0; i < n; i++) {
    const DMsg m = list[i];
    op_marg(tm, P, w.rd(P, m.proj_src), m.proj_src, w.tmp);
    op_absorb_ratio(tm, P, w, m.proj_dst, w.tmp, w.msg + m.slot);
    for (int k = tm.tid(); k < m.size; k += tm.size()) w.msg[m.slot + k] = w.tmp[k];
    tm.sync();
  }
}

// returns how many observations were entered
template <class Team>
__device__ int enter_row(Team& tm, const DProgram& P, Work& w, const int* obs, int n_obs,
                         const int* obs_proj) {
  int n = 0;
  for (int k = 0; k < n_obs; k++) {
    const int pj = obs_proj[k], o = obs[k];
    if (pj >= 0 && o >= 0) { op_evidence(tm, P, w, pj, o); n++; }
  }
  return n;
}

template <class Team>
__device__ void write_queries(Team& tm, const DProgram& P, Work& w, const DQuery& Q, double* row) {
  for (int q = 0; q < Q.n_query; q++) {
    const int pj = Q.proj[q], m = P.projs[pj].m;
    op_marg(tm, P, w.rd(P, pj), pj, w.scr);
    vec_normalise(tm, w.scr, m);
    for (int i = tm.tid(); i < m; i += tm.size()) row[Q.off[q] + i] = w.scr[i];
    tm.sync();
  }
}

// What a team needs besides the program: its workspace and the grid team's buffers.
struct TeamMem {
  double* gwork;
  size_t wstride;
  double* part;
  double* scratch;
  unsigned long long* trace;
  int seq_first, seq_step;
};

template <class Team>
__device__ __forceinline__ void attach_trace(Team&, const TeamMem&) {}
__device__ __forceinline__ void attach_trace(GridTeam& tm, const TeamMem& M) {
  tm.trace = M.trace;
  tm.seq_first = M.seq_first;
  tm.seq_step = M.seq_step;
}

template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_forward(DProgram P, DBatch B, DQuery Q, TeamMem M, int want_ll, int emit,
                             double* alpha, double* post, double* ll_out, int* status_out) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, tm.work(M.gwork, smem, M.wstride));
  double* vprev = w.va;
  double* vcur = w.va + P.S;
  for (int seq = tm.first(); seq < B.n_series; seq += tm.step()) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    double ll = 0;
    int bad = 0;
    for (int t = 0; t < T; t++) {
      load_tables(tm, P, w, t == 0 ? P.base0 : P.base1);
      double m1 = 0;
      if (t > 0 && P.nif > 0) op_absorb(tm, P, w, P.proj_in, vprev);
      if (want_ll) {
        if (t == 0) m1 = *P.m1_0;
        else if (P.nif == 0) m1 = P.R1[0];
        else {
          double s = 0;
          for (int i = tm.tid(); i < P.S; i += tm.size()) s += vprev[i] * P.R1[i];
          m1 = tm.sum(s);
        }
      }
      const int entered = enter_row(tm, P, w, B.obs + (row0 + t) * B.n_obs, B.n_obs, B.obs_proj);
      do_collect(tm, P, w);
      double m2 = 0;
      if (want_ll) m2 = vec_sum(tm, w.rd_clique(0) + P.root_tab, P.root_size);
      // a slice without any evidence has m2 == m1 by definition; do not let rounding decide
      // whether the running log-likelihood is "> 0" (the reference's BAD_LUCK test)
      if (want_ll && entered == 0) m2 = m1;
      if (emit) {
        do_distribute(tm, P, w, P.distribute, P.n_distribute);
        if (post) write_queries(tm, P, w, Q, post + (row0 + t) * Q.row);
      } else if (P.nif > 0)
        do_distribute(tm, P, w, P.path, P.n_path);
      if (P.nif > 0) {
        op_marg(tm, P, w.rd(P, P.proj_out), P.proj_out, vcur);
        vec_normalise(tm, vcur, P.S);
        if (alpha)
          for (int i = tm.tid(); i < P.S; i += tm.size()) alpha[(row0 + t) * P.S + i] = vcur[i];
        double* x = vprev; vprev = vcur; vcur = x;
        tm.sync();
      }
      if (want_ll) {  // src/nip.c:1458-1474 and the BAD_LUCK test of e_step, :1827-1831
        if (m1 > 0 && m2 > 0) ll += log(m2) - log(m1);
        if (m2 == 0) ll = -DBL_MAX;
        if (m1 <= 0 || m2 <= 0 || ll > 0) bad = 1;
      }
    }
    if (tm.tid() == 0) {
      if (ll_out) ll_out[seq] = ll;
      if (status_out) status_out[seq] = bad;
    }
  }
}

template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_backward(DProgram P, DBatch B, DQuery Q, TeamMem M, const double* alpha,
                              double* post, double* acc, long long acc_stride) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, tm.work(M.gwork, smem, M.wstride));
  double* a_prev = w.va;          // alpha_{t-1}
  double* a_cur = w.va + P.S;     // alpha_t
  double* gam = w.va + 2 * P.S;   // gamma_{t+1}
  double* my_acc = acc ? acc + (size_t)tm.slot() * acc_stride : nullptr;
  for (int seq = tm.first(); seq < B.n_series; seq += tm.step()) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    for (int t = T - 1; t >= 0; t--) {
      load_tables(tm, P, w, t == 0 ? P.base0 : P.base1);
      if (t > 0 && P.nif > 0) {
        for (int i = tm.tid(); i < P.S; i += tm.size()) a_prev[i] = alpha[(row0 + t - 1) * P.S + i];
        tm.sync();
        op_absorb(tm, P, w, P.proj_in, a_prev);
      }
      enter_row(tm, P, w, B.obs + (row0 + t) * B.n_obs, B.n_obs, B.obs_proj);
      if (t < T - 1 && P.nif > 0) op_absorb_ratio(tm, P, w, P.proj_out, gam, a_cur);
      do_collect(tm, P, w);
      do_distribute(tm, P, w, P.distribute, P.n_distribute);
      if (post) write_queries(tm, P, w, Q, post + (row0 + t) * Q.row);
      if (my_acc) {  // e_step "THE CORE", src/nip.c:1925-1967
        for (int v = 0; v < P.nv; v++) {
          if (t > 0 && (P.var_flags[v] & NIPGPU_IF_OLD_OUTGOING)) continue;
          const int pj = P.proj_fam[v], m = P.projs[pj].m;
          op_marg(tm, P, w.rd(P, pj), pj, w.scr);
          const double tot = vec_sum(tm, w.scr, m);
          if (tot != 0)
            for (int i = tm.tid(); i < m; i += tm.size()) my_acc[P.coff[v] + i] += w.scr[i] / tot;
          tm.sync();
        }
      }
      if (t > 0 && P.nif > 0) {
        op_marg(tm, P, w.rd(P, P.proj_in), P.proj_in, gam);
        vec_normalise(tm, gam, P.S);
      }
      double* x = a_prev; a_prev = a_cur; a_cur = x;
      tm.sync();
    }
  }
}

template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_likelihood(DProgram P, DBatch B, const int* proj_off, const int* proj_on,
                                TeamMem M, double* out) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, tm.work(M.gwork, smem, M.wstride));
  for (int seq = tm.first(); seq < B.n_series; seq += tm.step()) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    for (int t = 0; t < T; t++) {
      const int* obs = B.obs + (row0 + t) * B.n_obs;
      load_tables(tm, P, w, t == 0 ? P.base0 : P.base1);
      enter_row(tm, P, w, obs, B.n_obs, proj_off);
      do_collect(tm, P, w);
      const double m1 = vec_sum(tm, w.rd_clique(0) + P.root_tab, P.root_size);
      load_tables(tm, P, w, t == 0 ? P.base0 : P.base1);
      enter_row(tm, P, w, obs, B.n_obs, proj_off);
      const int extra = enter_row(tm, P, w, obs, B.n_obs, proj_on);
      do_collect(tm, P, w);
      double m2 = vec_sum(tm, w.rd_clique(0) + P.root_tab, P.root_size);
      if (extra == 0) m2 = m1;
      if (tm.tid() == 0) { out[(row0 + t) * 2] = m1; out[(row0 + t) * 2 + 1] = m2; }
    }
  }
}

// R1[i] = sum over everything but I_{t-1} of base1, m1_0 = total mass of base0
template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_calibrate(DProgram P, TeamMem M, double* R1, double* m1_0) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, M.gwork ? M.gwork : smem);
  load_tables(tm, P, w, P.base1);
  do_collect(tm, P, w);
  if (P.nif > 0) {
    do_distribute(tm, P, w, P.distribute, P.n_distribute);
    op_marg(tm, P, w.rd(P, P.proj_in), P.proj_in, w.va);
    for (int i = tm.tid(); i < P.S; i += tm.size()) R1[i] = w.va[i];
    tm.sync();
  } else {
    const double s = vec_sum(tm, w.rd_clique(0) + P.root_tab, P.root_size);
    if (tm.tid() == 0) R1[0] = s;
  }
  load_tables(tm, P, w, P.base0);
  do_collect(tm, P, w);
  const double s0 = vec_sum(tm, w.rd_clique(0) + P.root_tab, P.root_size);
  if (tm.tid() == 0) *m1_0 = s0;
}

template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_slice(DProgram P, TeamMem M, const double* start, double* out_tables,
                           double* out_msgs) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, M.gwork ? M.gwork : smem);
  load_tables(tm, P, w, start, false);
  do_collect(tm, P, w);
  do_distribute(tm, P, w, P.distribute, P.n_distribute);
  for (int i = tm.tid(); i < P.tab_total; i += tm.size()) out_tables[i] = w.tab[i];
  for (int i = tm.tid(); i < P.msg_total; i += tm.size()) out_msgs[i] = w.msg[i];
}

// make_consistent (src/nip.c:1600-1617) on WHATEVER state the caller holds: `in` = every
// clique's current table followed by every sepset's current potential (`new`), exactly what
// nip_collect_evidence / nip_distribute_evidence start from.  Each message pass is
// nip_message_pass (src/nipjointree.c:676-709): swap old/new, new = marginal of the sender,
// receiver *= new/old (x/0 -> 0).  `out` = consistent tables, then every sepset's `new` (the
// distribute message) and `old` (the collect message) — one contiguous block, so the host needs
// one copy in and one copy out per call.
template <class Team>
__global__ void __launch_bounds__(Team::kMaxThreads, Team::kMinBlocks) k_jt_propagate(DProgram P, TeamMem M, const double* in, double* out) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  Team tm(red, M.part, M.scratch, smem);
  attach_trace(tm, M);
  Work w = carve(P, M.gwork ? M.gwork : smem);
  load_tables(tm, P, w, in, false);
  for (int i = tm.tid(); i < P.msg_total; i += tm.size()) w.msg[i] = in[P.tab_total + i];
  tm.sync();
  for (int i = 0; i < P.n_collect; i++) {   // collect: old = what the sepset held, new = marginal
    const DMsg m = P.collect[i];
    op_marg(tm, P, w.tab, m.proj_src, w.tmp);
    op_absorb_ratio(tm, P, w, m.proj_dst, w.tmp, w.msg + m.slot);
    for (int k = tm.tid(); k < m.size; k += tm.size()) w.msg[m.slot + k] = w.tmp[k];
    tm.sync();
  }
  double* out_new = out + P.tab_total;
  double* out_old = out_new + P.msg_total;
  for (int i = tm.tid(); i < P.msg_total; i += tm.size()) out_old[i] = w.msg[i];
  tm.sync();
  do_distribute(tm, P, w, P.distribute, P.n_distribute);
  for (int i = tm.tid(); i < P.tab_total; i += tm.size()) out[i] = w.tab[i];
  for (int i = tm.tid(); i < P.msg_total; i += tm.size()) out_new[i] = w.msg[i];
}

// single-slice API: mass = sum of cliques - sum of sepsets (nip_probability_mass,
// src/nipjointree.c:1156-1188) of the consistent tables left by k_jt_slice
__global__ void k_jt_mass(const double* tables, int n_tab, const double* msgs, int n_msg, double* out) {
  __shared__ double red[40];
  CtaTeam tm(red, nullptr, nullptr, nullptr);
  const double a = vec_sum(tm, tables, n_tab);
  const double b = vec_sum(tm, msgs, n_msg);
  if (threadIdx.x == 0) *out = a - b;
}

// get_probability (src/nip.c:2261-2298): normalised one-variable marginal of the family clique
__global__ void k_jt_marginal(DProgram P, const double* tables, int pj, double* out) {
  __shared__ double red[40];
  CtaTeam tm(red, nullptr, nullptr, nullptr);
  const int m = P.projs[pj].m;
  op_marg(tm, P, tables, pj, out);
  vec_normalise(tm, out, m);
}

TeamMem team_mem(const DProgram& p, const JtLaunch& l) {
  TeamMem M;
  M.gwork = l.gwork;
  M.wstride = jt_work_doubles(p, l.mode == JT_MODE_GRID);
  M.part = l.part;
  M.scratch = l.scratch;
  M.trace = l.trace;
  M.seq_first = 0;
  M.seq_step = 1;
  return M;
}

// memory of group g of a grid-mode launch
TeamMem group_mem(const DProgram& p, const JtLaunch& l, int g, int groups) {
  TeamMem M = team_mem(p, l);
  M.gwork = l.gwork + (size_t)g * l.group_stride;
  M.part = l.part + (size_t)g * l.group_stride;
  M.scratch = l.scratch + (size_t)g * l.group_stride;
  if (g > 0) M.trace = nullptr;
  M.seq_first = g;
  M.seq_step = groups;
  return M;
}

// grid mode: the groups' cooperative kernels side by side (group 0 on the caller's stream)
template <class Launch>
int launch_groups(const JtLaunch& l, int groups, cudaStream_t st, Launch launch) {
  if (groups <= 1) return launch(0, 1, st);
  NIPGPU_CUDA(cudaEventRecord(l.aux_event[0], st));
  for (int g = 1; g < groups; g++) {
    NIPGPU_CUDA(cudaStreamWaitEvent(l.aux_stream[g - 1], l.aux_event[0], 0));
    if (int e = launch(g, groups, l.aux_stream[g - 1])) return e;
    NIPGPU_CUDA(cudaEventRecord(l.aux_event[g], l.aux_stream[g - 1]));
  }
  if (int e = launch(0, groups, st)) return e;
  for (int g = 1; g < groups; g++) NIPGPU_CUDA(cudaStreamWaitEvent(st, l.aux_event[g], 0));
  return NIPGPU_OK;
}

// One launcher for the three instantiations of a kernel template.
template <class KW, class KC, class KG, class... Args>
int launch_team(KW kw, KC kc, KG kg, const JtLaunch& l, int grid, cudaStream_t st, Args... args) {
  if (l.mode == JT_MODE_GRID) {
    void* argv[] = {(void*)&args...};
    if (l.smem_bytes > 48 * 1024)
      NIPGPU_CUDA(cudaFuncSetAttribute(kg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes));
    NIPGPU_CUDA(cudaLaunchCooperativeKernel((const void*)kg, dim3(grid), dim3(l.threads), argv, l.smem_bytes, st));
  } else if (l.mode == JT_MODE_WARP) {
    if (l.smem_bytes > 48 * 1024)
      NIPGPU_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes));
    kw<<<grid, l.threads, l.smem_bytes, st>>>(args...);
  } else {
    if (l.smem_bytes > 48 * 1024)
      NIPGPU_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes));
    kc<<<grid, l.threads, l.smem_bytes, st>>>(args...);
  }
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

}  // namespace

int jt_grid_ctas(int threads, int sm_count, size_t* smem_bytes) {
  const size_t smem = (size_t)kRingSlots * threads * sizeof(double2);
  *smem_bytes = smem;
  int per_sm = 1 << 30;
  auto probe = [&](auto kernel) {
    int n = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    per_sm = std::min(per_sm, n);
  };
  probe(k_jt_forward<GridTeam>);
  probe(k_jt_backward<GridTeam>);
  probe(k_jt_likelihood<GridTeam>);
  probe(k_jt_calibrate<GridTeam>);
  probe(k_jt_slice<GridTeam>);
  probe(k_jt_propagate<GridTeam>);
  return sm_count * per_sm;
}

JtLaunch jt_fit(const JtLaunch& l, int n_series) {
  JtLaunch r = l;
  const int n = std::max(n_series, 1);
  if (l.mode == JT_MODE_WARP) {
    const int wpc = l.threads / 32;
    r.grid = std::max(1, std::min(l.grid, (n + wpc - 1) / wpc));
    r.slots = r.grid * wpc;
  } else if (l.mode == JT_MODE_GRID) {
    r.groups = std::max(1, std::min(l.groups, n));
    r.slots = r.groups;
  } else {
    r.grid = std::max(1, std::min(l.grid, n));
    r.slots = r.grid;
  }
  return r;
}

int jt_forward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l, int want_ll,
               int emit, double* alpha, double* post, double* ll, int* status, cudaStream_t st) {
  if (l.mode == JT_MODE_GRID)
    return launch_groups(l, l.groups, st, [&](int g, int groups, cudaStream_t s) {
      return launch_team(k_jt_forward<WarpTeam>, k_jt_forward<CtaTeam>, k_jt_forward<GridTeam>, l, l.grid, s,
                         p, b, q, group_mem(p, l, g, groups), want_ll, emit, alpha, post, ll, status);
    });
  return launch_team(k_jt_forward<WarpTeam>, k_jt_forward<CtaTeam>, k_jt_forward<GridTeam>, l, l.grid, st,
                     p, b, q, team_mem(p, l), want_ll, emit, alpha, post, ll, status);
}

int jt_backward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l,
                const double* alpha, double* post, double* acc, long long acc_stride, cudaStream_t st) {
  if (l.mode == JT_MODE_GRID)
    return launch_groups(l, l.groups, st, [&](int g, int groups, cudaStream_t s) {
      return launch_team(k_jt_backward<WarpTeam>, k_jt_backward<CtaTeam>, k_jt_backward<GridTeam>, l, l.grid,
                         s, p, b, q, group_mem(p, l, g, groups), alpha, post, acc, acc_stride);
    });
  return launch_team(k_jt_backward<WarpTeam>, k_jt_backward<CtaTeam>, k_jt_backward<GridTeam>, l, l.grid,
                     st, p, b, q, team_mem(p, l), alpha, post, acc, acc_stride);
}

int jt_likelihood(const DProgram& p, const DBatch& b, const int* proj_off, const int* proj_on,
                  const JtLaunch& l, double* out, cudaStream_t st) {
  if (l.mode == JT_MODE_GRID)
    return launch_groups(l, l.groups, st, [&](int g, int groups, cudaStream_t s) {
      return launch_team(k_jt_likelihood<WarpTeam>, k_jt_likelihood<CtaTeam>, k_jt_likelihood<GridTeam>, l,
                         l.grid, s, p, b, proj_off, proj_on, group_mem(p, l, g, groups), out);
    });
  return launch_team(k_jt_likelihood<WarpTeam>, k_jt_likelihood<CtaTeam>, k_jt_likelihood<GridTeam>, l,
                     l.grid, st, p, b, proj_off, proj_on, team_mem(p, l), out);
}

// single-sequence kernels run one team; in warp mode that is a CTA of one warp
static JtLaunch single_team(const JtLaunch& l, const DProgram& p) {
  JtLaunch one = l;
  if (l.mode == JT_MODE_WARP) {
    one.mode = JT_MODE_CTA;
    one.threads = 32;
    one.smem_bytes = jt_work_doubles(p) * sizeof(double);
  }
  return one;
}

int jt_calibrate(const DProgram& p, const JtLaunch& l, double* R1, double* m1_0, cudaStream_t st) {
  const JtLaunch one = single_team(l, p);
  return launch_team(k_jt_calibrate<CtaTeam>, k_jt_calibrate<CtaTeam>, k_jt_calibrate<GridTeam>, one,
                     one.mode == JT_MODE_GRID ? one.grid : 1, st, p, team_mem(p, one), R1, m1_0);
}

// ---- memoised likelihood loop -------------------------------------------------------------
// The slices of the likelihood loop are independent, so (m1, m2) of a record depends only on
// its evidence configuration and on whether it opens a series.  When the configurations are
// few, k_jt_likelihood runs once per configuration and the records only gather.
__global__ void k_jt_first_rows(const long long* row_off, int n_series, long long rows, unsigned char* first) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_series && row_off[s] < rows) first[row_off[s]] = 1;
}

__global__ void k_jt_like_gather(const int* __restrict__ obs, int n_obs, long long rows,
                                 const unsigned char* __restrict__ first, const int* __restrict__ col_stride,
                                 const int* __restrict__ col_card, const double2* __restrict__ table,
                                 double2* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    int cfg = 0;
    for (int k = 0; k < n_obs; k++) {
      const int st = col_stride[k];
      if (st == 0) continue;  // column outside both evidence sets
      const int o = obs[r * n_obs + k], card = col_card[k];
      cfg += st * (o < 0 ? 0 : (o < card ? o + 1 : card + 1));  // missing / state / out of range
    }
    out[r] = table[2 * cfg + (first[r] ? 0 : 1)];
  }
}

int jt_first_rows(const long long* row_off, int n_series, long long rows, unsigned char* first, cudaStream_t st) {
  if (n_series <= 0) return NIPGPU_OK;
  k_jt_first_rows<<<(n_series + 255) / 256, 256, 0, st>>>(row_off, n_series, rows, first);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

// one data column (niplikelihood over single-variable records, config C5): four consecutive records
// per thread and step — one 16-byte load of the observations, one 4-byte load of the series-start
// flags, two 32-byte stores — so that enough bytes are in flight per SM to approach the HBM rate
__global__ void __launch_bounds__(256) k_jt_like_gather1(const int* __restrict__ obs, long long rows,
                                                         const unsigned char* __restrict__ first, int stride,
                                                         int card, const double2* __restrict__ table,
                                                         double* __restrict__ out) {
  const long long quads = rows >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    const int4 o = __ldg(reinterpret_cast<const int4*>(obs) + q);
    const uchar4 f = __ldg(reinterpret_cast<const uchar4*>(first) + q);
    auto look = [&](int ob, unsigned char fr) {
      const int cfg = stride * (ob < 0 ? 0 : (ob < card ? ob + 1 : card + 1));
      return table[2 * cfg + (fr ? 0 : 1)];
    };
    const double2 a = look(o.x, f.x), b = look(o.y, f.y), c = look(o.z, f.z), d = look(o.w, f.w);
    double* dst = out + 8 * q;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4), "d"(c.x), "d"(c.y), "d"(d.x), "d"(d.y) : "memory");
  }
  // the last rows % 4 records
  const long long r = 4 * quads + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows) {
    const int ob = obs[r];
    const int cfg = stride * (ob < 0 ? 0 : (ob < card ? ob + 1 : card + 1));
    reinterpret_cast<double2*>(out)[r] = table[2 * cfg + (first[r] ? 0 : 1)];
  }
}

int jt_like_gather(const int* obs, int n_obs, long long rows, const unsigned char* first, const int* col_stride,
                   const int* col_card, const double* table, double* out, int sm_count, cudaStream_t st,
                   int host_stride0, int host_card0) {
  if (rows <= 0) return NIPGPU_OK;
  if (n_obs == 1 && host_stride0 > 0 && ((uintptr_t)obs % 16 == 0) && ((uintptr_t)first % 4 == 0) &&
      ((uintptr_t)out % 32 == 0)) {
    const long long want = ((rows >> 2) + 255) / 256;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * 8));
    k_jt_like_gather1<<<grid, 256, 0, st>>>(obs, rows, first, host_stride0, host_card0,
                                            reinterpret_cast<const double2*>(table), out);
    NIPGPU_LAUNCHED();
    return NIPGPU_OK;
  }
  const long long want = (rows + 255) / 256;
  const int grid = (int)std::min<long long>(want, (long long)sm_count * 16);
  k_jt_like_gather<<<grid, 256, 0, st>>>(obs, n_obs, rows, first, col_stride, col_card,
                                         reinterpret_cast<const double2*>(table), reinterpret_cast<double2*>(out));
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_mass(const double* tables, int n_tab, const double* msgs, int n_msg, double* out, cudaStream_t st) {
  k_jt_mass<<<1, 256, 0, st>>>(tables, n_tab, msgs, n_msg, out);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_marginal(const DProgram& p, const double* tables, int proj, double* out, cudaStream_t st) {
  k_jt_marginal<<<1, 128, 0, st>>>(p, tables, proj, out);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_slice(const DProgram& p, const JtLaunch& l, const double* start, double* out_tables,
             double* out_msgs, cudaStream_t st) {
  const JtLaunch one = single_team(l, p);
  return launch_team(k_jt_slice<CtaTeam>, k_jt_slice<CtaTeam>, k_jt_slice<GridTeam>, one,
                     one.mode == JT_MODE_GRID ? one.grid : 1, st, p, team_mem(p, one), start, out_tables,
                     out_msgs);
}

int jt_propagate(const DProgram& p, const JtLaunch& l, const double* in, double* out, cudaStream_t st) {
  const JtLaunch one = single_team(l, p);
  return launch_team(k_jt_propagate<CtaTeam>, k_jt_propagate<CtaTeam>, k_jt_propagate<GridTeam>, one,
                     one.mode == JT_MODE_GRID ? one.grid : 1, st, p, team_mem(p, one), in, out);
}

}  // namespace nipgpu
