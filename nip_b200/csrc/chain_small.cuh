// chain_small.cuh — the chain engine for interfaces of at most 8 joint states (configs C1 / C5:
// examples/model.net, 4 states) on LARGE batches: one THREAD per sequence.
//
// For such a model a slice costs a few dozen FP64 operations and moves
//     32 B forward row (write) + 32 B (read back) + 32 B posterior + 2 x 4 B evidence index
// so the pass is bound by HBM, not by the tensor pipe: the DMMA kernels (8 sequences per
// warp, one 8-state tile) leave most of every tile and most of the memory system idle.  Here a
// warp carries 32 sequences, the S x S transition table sits in registers / shared memory, and
// the two arrays that are private to the engine — forward rows and evidence indices — are kept
// TIME-MAJOR over the length-sorted sequences
//     element (t, p)  at  toff[t] + p,   toff[t] = number of (slice, sequence) pairs before slice t
// (sequences sorted by length, longest first: the ones alive at slice t are p < n_t), so a warp
// reads and writes them as whole 128-byte lines.  Only the posterior rows, whose layout belongs
// to the caller (series after series), are written 32 bytes per thread.
//
// Arithmetic: the literal scaled recursions of src/nip.c:1103-1315 / 1320-1581 for this model
// class (see chain.cuh): alpha normalised every slice (zero sum: left as it is,
// nip_normalise_array), m1_t = alpha_{t-1} . R1, m2_t = sum of the unnormalised alpha_t, the
// backward pass as beta_{t-1} ~ A (lambda_t * beta_t), posterior = normalise(alpha_t * beta_t).
// Included by chain.cu (inside its anonymous namespace).

// running log-likelihood: LogAcc (chain.cu)

template <int S>
struct SmallGeom {
  static constexpr int AS = S + (S & 1);          // doubles per forward row (16-byte units)
  static constexpr bool V4 = AS % 4 == 0;         // rows are multiples of 32 bytes
};

template <int N>
__device__ __forceinline__ void small_store_row(double* p, const double (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "d"(v[i]), "d"(v[i + 1]), "d"(v[i + 2]),
                   "d"(v[i + 3])
                   : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v[i], v[i + 1]);
  }
}

template <int N>
__device__ __forceinline__ void small_load_row(const double* p, double (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(v[i]), "=d"(v[i + 1]), "=d"(v[i + 2]), "=d"(v[i + 3])
                   : "l"(p + i));
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const double2 x = *reinterpret_cast<const double2*>(p + i);
      v[i] = x.x;
      v[i + 1] = x.y;
    }
  }
}

// posterior / filtered row of the caller: S doubles at any 8-byte aligned address
template <int S>
__device__ __forceinline__ void small_store_post(double* p, const double (&v)[S + (S & 1)], bool wide) {
  if (wide) {
    small_store_row<S + (S & 1)>(p, v);   // only taken when S is even and the row is 16/32-byte aligned
  } else {
#pragma unroll
    for (int i = 0; i < S; i++) p[i] = v[i];
  }
}

struct SmallDev {
  int S, SP, c_miss, n_comb;
  double m1_0;
  const double *A;        // [S][S] row = previous state, column = current state (base1)
  const double *phi0, *R1, *lam_comb;   // [SP], [SP], [n_comb][SP]
  const long long* toff;  // [t_max + 1]
  int n_toff;             // entries of toff staged in shared memory (0: read from global)
  const int* cfgT;        // time-major evidence index
  int n_series;
  const int* order;       // sorted position -> series
  const int* len_sorted;
  const long long* row_off;
};

// shared memory: A [S*S] | lam_comb [n_lam][S] (when it fits: n_lam = n_comb, else 0) | toff [n_toff]
// (the time-major offsets, when they fit: otherwise every slice's address waits for a global load)
template <int S, bool FILT, bool WLL>
__global__ void __launch_bounds__(128) k_chain_small_forward(SmallDev C, int n_lam, int store_alpha,
                                                             double* __restrict__ alphaT,
                                                             double* __restrict__ post, int post_stride,
                                                             int post_off, int post_wide, double* ll_out,
                                                             int* status_out) {
  constexpr int AS = SmallGeom<S>::AS;
  extern __shared__ double sm[];
  double* sA = sm;
  double* sL = sm + S * S;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) sA[i] = C.A[i];
  for (int i = threadIdx.x; i < n_lam * S; i += blockDim.x) sL[i] = C.lam_comb[(long long)(i / S) * C.SP + i % S];
  long long* sT = reinterpret_cast<long long*>(sL + n_lam * S);
  for (int i = threadIdx.x; i < C.n_toff; i += blockDim.x) sT[i] = C.toff[i];
  const long long* toff = C.n_toff ? sT : C.toff;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C.n_series) return;
  const int T = C.len_sorted[p];
  if (T <= 0) {
    if (ll_out) ll_out[C.order[p]] = 0.0;
    if (status_out) status_out[C.order[p]] = 0;
    return;
  }
  const int orig = C.order[p];
  const long long row0 = C.row_off[orig];
  // the transition table: registers up to 4 states, shared-memory broadcasts above
  constexpr int AR = S <= 4 ? S : 1;
  double Areg[AR][AR];
#pragma unroll
  for (int i = 0; i < AR; i++)
#pragma unroll
    for (int j = 0; j < AR; j++) Areg[i][j] = sA[i * S + j];
  auto A = [&](int i, int j) -> double {
    if constexpr (S <= 4) return Areg[i][j];
    else return sA[i * S + j];
  };
  double r1[S];
#pragma unroll
  for (int i = 0; i < S; i++) r1[i] = WLL ? C.R1[i] : 0.0;

  auto lam_row = [&](int c, double (&lam)[S]) {
    if (n_lam) {
#pragma unroll
      for (int i = 0; i < S; i++) lam[i] = sL[c * S + i];
    } else {
#pragma unroll
      for (int i = 0; i < S; i++) lam[i] = __ldg(C.lam_comb + (long long)c * C.SP + i);
    }
  };
  LogAcc L;
  double a[AS];
  if constexpr (AS > S) a[S] = 0.0;
  int c = __ldg(C.cfgT + toff[0] + p);
  int c_next = T > 1 ? __ldg(C.cfgT + toff[1] + p) : 0;
  {  // slice 0: alpha_0 = normalise(phi0 * lambda_0), m1 = mass of the evidence-free slice
    double lam[S], s = 0;
    lam_row(c, lam);
#pragma unroll
    for (int i = 0; i < S; i++) { a[i] = C.phi0[i] * lam[i]; s += a[i]; }
    if (WLL) L.add(C.m1_0, c == C.c_miss ? C.m1_0 : s, true);
    const double inv = safe_rcp(s);
#pragma unroll
    for (int i = 0; i < S; i++) a[i] *= inv;
  }
  for (int t = 0;; t++) {
    if (store_alpha) small_store_row<AS>(alphaT + (toff[t] + p) * AS, a);
    if (FILT) small_store_post<S>(post + (row0 + t) * post_stride + post_off, a, post_wide);
    if (t + 1 >= T) break;
    c = c_next;
    if (t + 2 < T) c_next = __ldg(C.cfgT + toff[t + 2] + p);
    double lam[S], u[S], s = 0, m1 = 0;
    lam_row(c, lam);
#pragma unroll
    for (int j = 0; j < S; j++) u[j] = a[0] * A(0, j);
#pragma unroll
    for (int i = 1; i < S; i++)
#pragma unroll
      for (int j = 0; j < S; j++) u[j] = fma(a[i], A(i, j), u[j]);
    if (WLL) {
#pragma unroll
      for (int i = 0; i < S; i++) m1 = fma(a[i], r1[i], m1);
    }
#pragma unroll
    for (int j = 0; j < S; j++) { u[j] *= lam[j]; s += u[j]; }
    if (WLL) L.add(m1, c == C.c_miss ? m1 : s, true);
    const double inv = safe_rcp(s);
#pragma unroll
    for (int j = 0; j < S; j++) a[j] = u[j] * inv;
  }
  if (ll_out) ll_out[orig] = WLL ? L.value() : 0.0;
  if (status_out) status_out[orig] = WLL ? L.bad : 0;
}

template <int S>
__global__ void __launch_bounds__(128) k_chain_small_backward(SmallDev C, int n_lam, const double* __restrict__ alphaT,
                                                              double* __restrict__ post, int post_stride,
                                                              int post_off, int post_wide) {
  constexpr int AS = SmallGeom<S>::AS;
  extern __shared__ double sm[];
  double* sA = sm;
  double* sL = sm + S * S;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) sA[i] = C.A[i];
  for (int i = threadIdx.x; i < n_lam * S; i += blockDim.x) sL[i] = C.lam_comb[(long long)(i / S) * C.SP + i % S];
  long long* sT = reinterpret_cast<long long*>(sL + n_lam * S);
  for (int i = threadIdx.x; i < C.n_toff; i += blockDim.x) sT[i] = C.toff[i];
  const long long* toff = C.n_toff ? sT : C.toff;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C.n_series) return;
  const int T = C.len_sorted[p];
  if (T <= 0) return;
  const long long row0 = C.row_off[C.order[p]];
  // the transition table: registers up to 4 states, shared-memory broadcasts above
  constexpr int AR = S <= 4 ? S : 1;
  double Areg[AR][AR];
#pragma unroll
  for (int i = 0; i < AR; i++)
#pragma unroll
    for (int j = 0; j < AR; j++) Areg[i][j] = sA[i * S + j];
  auto A = [&](int i, int j) -> double {
    if constexpr (S <= 4) return Areg[i][j];
    else return sA[i * S + j];
  };
  auto lam_row = [&](int c, double (&lam)[S]) {
    if (n_lam) {
#pragma unroll
      for (int i = 0; i < S; i++) lam[i] = sL[c * S + i];
    } else {
#pragma unroll
      for (int i = 0; i < S; i++) lam[i] = __ldg(C.lam_comb + (long long)c * C.SP + i);
    }
  };
  double beta[S], a[AS], an[AS];
#pragma unroll
  for (int i = 0; i < S; i++) beta[i] = 1.0;
  small_load_row<AS>(alphaT + (toff[T - 1] + p) * AS, a);
  int c = __ldg(C.cfgT + toff[T - 1] + p);
  for (int t = T - 1;; t--) {
    // requests of the next iteration first: they do not depend on the recursion
    int c_prev = 0;
    if (t >= 1) {
      small_load_row<AS>(alphaT + (toff[t - 1] + p) * AS, an);
      c_prev = __ldg(C.cfgT + toff[t - 1] + p);
    }
    double g[AS], s = 0;
    if constexpr (AS > S) g[S] = 0.0;
#pragma unroll
    for (int i = 0; i < S; i++) { g[i] = a[i] * beta[i]; s += g[i]; }
    const double inv = safe_rcp(s);
#pragma unroll
    for (int i = 0; i < S; i++) g[i] *= inv;
    small_store_post<S>(post + (row0 + t) * post_stride + post_off, g, post_wide);
    if (t == 0) break;
    // beta_{t-1} = A (lambda_t * beta_t), scaled by the reciprocal of its sum
    double lam[S], r[S], u[S], bs = 0;
    lam_row(c, lam);
#pragma unroll
    for (int j = 0; j < S; j++) r[j] = lam[j] * beta[j];
#pragma unroll
    for (int i = 0; i < S; i++) {
      u[i] = A(i, 0) * r[0];
#pragma unroll
      for (int j = 1; j < S; j++) u[i] = fma(A(i, j), r[j], u[i]);
      bs += u[i];
    }
    const double binv = safe_rcp(bs);
#pragma unroll
    for (int i = 0; i < S; i++) beta[i] = u[i] * binv;
#pragma unroll
    for (int i = 0; i < AS; i++) a[i] = an[i];
    c = c_prev;
  }
}

// time-major copy of the per-row evidence index (once per evidence plan)
__global__ void k_chain_small_cfgT(const int* __restrict__ cfg, const long long* __restrict__ toff,
                                   const long long* __restrict__ row_off, const int* __restrict__ order,
                                   const int* __restrict__ len_sorted, int n_series, int* __restrict__ cfgT) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_series) return;
  const int T = len_sorted[p];
  const int* src = cfg + row_off[order[p]];
  for (int t = 0; t < T; t++) cfgT[toff[t] + p] = src[t];
}

// A[i_prev][i_cur] = base1 entry of the interface clique
__global__ void k_chain_small_A(const double* Bf1, int S, double* A) {
  const int x = threadIdx.x;
  if (x < S * S) A[x] = Bf1[frag_index(x / S, x % S, 1)];
}
