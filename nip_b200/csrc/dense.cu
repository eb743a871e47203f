// dense.cu — chain engine for LARGE interfaces (|I| > 64, config C4: 1024 states).
//
// Same recursions as chain.cu (lagged scaling, scaled beta), but one slice of the
// whole batch at a time: the slice-to-slice contraction
//     Phi[B x S] = Own_{t-1}[B x S] . A[S x S]          (forward)
//     U  [B x S] = R_t     [B x S] . A^T[S x S]         (backward)
// is a real dense FP64 GEMM and runs as a shared-memory tiled DMMA kernel
// (mma.sync m8n8k4.f64, 128x128x16 tiles, cp.async double buffering) with the
// evidence row / scale factors applied in the epilogue.  The O(B.S) bookkeeping
// of a slice (masses, likelihood, normalisation of posteriors, scale factors) is
// a separate "settle" kernel, one warp per sequence, between two GEMMs.
//
// Reference semantics: start/finish_timeslice_message_pass (src/nip.c:1031-1098),
// forward_inference / forward_backward_inference (src/nip.c:1103-1581).
#include "chain.cuh"

#include <algorithm>
#include <cfloat>
#include <functional>

namespace nipgpu {
namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool on) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = on ? 16 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ double rcp_or_one(double x) { return x != 0 ? 1.0 / x : 1.0; }

struct DenseBatch {
  int n_series;            // sequences, sorted by length (descending)
  const int* order;        // sorted position -> series
  const int* len_sorted;
  const long long* row_off;
  const int* cfg;          // [rows] combined evidence index
};

// per-sequence running state of the forward pass (sorted order)
struct DenseState {
  double *K, *Pp, *cp, *g;          // see chain.cu: m2 = c*K; m1 numerator; previous c; scale of the slice
  double *p1, *p2;                  // LogAcc mantissa products
  int *e1, *e2, *zero, *bad, *noev;
};

constexpr int TM = 128, TN = 128, TK = 16, LDA = TK + 2, LDB = TN + 4;

// C[TM x TN] tile of X[n_rows x SP] . M[SP x SP]; X rows are addressed through `xrow(b)`.
// MODE 0 (forward):  out row b, col j = C * lam_comb[cfg(b,t)][j] * g[b]      -> alpha row t
// MODE 1 (backward): beta = C * h[b] -> Bout;  R = beta * lam_comb[cfg(b,t-1)][j] -> Rout
template <int MODE>
__global__ void __launch_bounds__(256) k_dense_gemm(DenseBatch B, int t, int n_rows, int SP,
                                                    const double* __restrict__ X, long long x_stride_is_alpha,
                                                    const double* __restrict__ M,
                                                    const double* __restrict__ lam_comb,
                                                    const double* __restrict__ scale,
                                                    double* __restrict__ out0, double* __restrict__ out1) {
  extern __shared__ double sm[];
  double* sA = sm;                         // [2][TM][LDA]
  double* sB = sm + 2 * TM * LDA;          // [2][TK][LDB]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  const int wm = w & 3, wn = w >> 2;       // warp tile: rows 32*wm.., cols 64*wn..
  const int row_base = blockIdx.y * TM, col_base = blockIdx.x * TN;
  // row pointers of this CTA's A tile: thread r < TM owns row r
  __shared__ const double* s_xrow[TM];
  __shared__ long long s_out_row[TM];
  __shared__ int s_cfg[TM];
  __shared__ double s_scale[TM];
  if (tid < TM) {
    const int b = row_base + tid;
    const bool on = b < n_rows;
    const long long r0 = on ? B.row_off[B.order[b]] : 0;
    if (x_stride_is_alpha) s_xrow[tid] = on ? X + (r0 + t - 1) * SP : nullptr;      // own_{t-1}
    else s_xrow[tid] = on ? X + (long long)b * SP : nullptr;                            // R_t (sorted order)
    s_out_row[tid] = on ? (MODE == 0 ? r0 + t : (long long)b) : -1;
    s_cfg[tid] = on ? B.cfg[r0 + (MODE == 0 ? t : t - 1)] : 0;
    s_scale[tid] = on ? scale[b] : 0.0;
  }
  __syncthreads();
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_tiles = [&](int buf, int k0) {
    // A tile: TM rows x TK doubles = TM*8 16-byte pieces; B tile: TK rows x TN doubles = TK*64 pieces
#pragma unroll
    for (int x = tid; x < TM * (TK / 2); x += 256) {
      const int r = x / (TK / 2), c = 2 * (x % (TK / 2));
      const double* src = s_xrow[r];
      cp_async16(sA + (buf * TM + r) * LDA + c, src ? src + k0 + c : X, src != nullptr);
    }
#pragma unroll
    for (int x = tid; x < TK * (TN / 2); x += 256) {
      const int r = x / (TN / 2), c = 2 * (x % (TN / 2));
      cp_async16(sB + (buf * TK + r) * LDB + c, M + (long long)(k0 + r) * SP + col_base + c, true);
    }
    cp_async_commit();
  };
  const int nk = SP / TK;
  load_tiles(0, 0);
  for (int kc = 0; kc < nk; kc++) {
    if (kc + 1 < nk) { load_tiles((kc + 1) & 1, (kc + 1) * TK); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* a_s = sA + ((kc & 1) * TM + 32 * wm) * LDA;
    const double* b_s = sB + (kc & 1) * TK * LDB + 64 * wn;
#pragma unroll
    for (int kk = 0; kk < TK / 4; kk++) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = a_s[(8 * i + g) * LDA + 4 * kk + q];
#pragma unroll
      for (int j = 0; j < 8; j++) bf[j] = b_s[(4 * kk + q) * LDB + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
  // epilogue: lane holds C[32*wm + 8*i + g][64*wn + 8*j + 2q + {0,1}]
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = 32 * wm + 8 * i + g;
    const long long orow = s_out_row[r];
    if (orow < 0) continue;
    const double sc = s_scale[r];
    const double* lrow = lam_comb + (long long)s_cfg[r] * SP + col_base + 64 * wn;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = 8 * j + 2 * q;
      const double2 l = *reinterpret_cast<const double2*>(lrow + c);
      const long long o = orow * SP + col_base + 64 * wn + c;
      if (MODE == 0) {
        *reinterpret_cast<double2*>(out0 + o) = make_double2(acc[i][j][0] * l.x * sc, acc[i][j][1] * l.y * sc);
      } else {
        const double b0 = acc[i][j][0] * sc, b1 = acc[i][j][1] * sc;
        *reinterpret_cast<double2*>(out0 + o) = make_double2(b0, b1);
        *reinterpret_cast<double2*>(out1 + o) = make_double2(b0 * l.x, b1 * l.y);
      }
    }
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void logacc_renorm(double& p, int& e) {
  const int hi = __double2hiint(p);
  const int ex = ((hi >> 20) & 0x7ff) - 1023;
  e += ex;
  p = __hiloint2double(hi - (ex << 20), __double2loint(p));
}

// slice 0 of every sequence: own_0 = phi0 * lambda_0 / S0; one warp per sequence
__global__ void k_dense_first(DenseBatch B, int S, int SP, const double* phi0, const double* lam_comb,
                              double m1_0, int c_miss, double* alpha, DenseState st) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B.n_series || B.len_sorted[b] < 1) return;
  const long long r0 = B.row_off[B.order[b]];
  const int c = B.cfg[r0];
  const double* l = lam_comb + (long long)c * SP;
  double s = 0;
  for (int j = lane; j < SP; j += 32) s += phi0[j] * l[j];
  s = warp_sum_d(s);
  const double inv = rcp_or_one(s);
  for (int j = lane; j < SP; j += 32) alpha[r0 * SP + j] = phi0[j] * l[j] * inv;
  if (lane == 0) {
    st.K[b] = s; st.Pp[b] = m1_0; st.cp[b] = 1.0; st.noev[b] = c == c_miss;
    st.p1[b] = 1.0; st.p2[b] = 1.0; st.e1[b] = 0; st.e2[b] = 0; st.zero[b] = 0; st.bad[b] = 0;
  }
  (void)S;
}

// settles slice t-1 of every sequence that has it (vector = alpha row t-1): masses, likelihood,
// filtered output, and the scale g of slice t.  One warp per sequence.
__global__ void k_dense_fsettle(DenseBatch B, int t, int S, int SP, const double* R1, int c_miss,
                                const double* alpha, int want_ll, double* post, int post_stride,
                                int post_off, DenseState st) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B.n_series || B.len_sorted[b] < t) return;    // slice t-1 must exist
  const long long r0 = B.row_off[B.order[b]];
  const double* own = alpha + (r0 + t - 1) * SP;
  double c = 0, P = 0;
  for (int j = lane; j < SP; j += 32) { const double x = own[j]; c += x; P += x * R1[j]; }
  c = warp_sum_d(c);
  P = warp_sum_d(P);
  const double m2p = c * st.K[b];
  if (post) {   // filtering: alpha_{t-1} = own / c
    const double cinv = rcp_or_one(c);
    double* prow = post + (r0 + t - 1) * post_stride + post_off;
    for (int j = lane; j < S; j += 32) prow[j] = own[j] * cinv;
  }
  if (lane == 0) {
    if (want_ll) {   // LogAcc::add of chain.cu (src/nip.c:1458-1474, BAD_LUCK test :1827-1831)
      const double m1 = st.Pp[b], m2 = st.noev[b] ? st.Pp[b] : m2p * st.cp[b];
      const bool both = m1 > 0 && m2 > 0;
      double p1 = st.p1[b] * (both ? m1 : 1.0), p2 = st.p2[b] * (both ? m2 : 1.0);
      int e1 = st.e1[b], e2 = st.e2[b];
      logacc_renorm(p1, e1);
      logacc_renorm(p2, e2);
      if (m2 == 0) st.zero[b] = 1;
      const bool pos = e2 > e1 || (e2 == e1 && p2 > p1);
      if (m1 <= 0 || m2 <= 0 || (pos && !st.zero[b])) st.bad[b] = 1;
      st.p1[b] = p1; st.p2[b] = p2; st.e1[b] = e1; st.e2[b] = e2;
    }
    const double den = c * m2p;
    st.g[b] = rcp_or_one(den);
    st.K[b] = den != 0 ? m2p : 0.0;
    st.Pp[b] = P;
    st.cp[b] = c;
    if (t < B.len_sorted[b]) st.noev[b] = B.cfg[r0 + t] == c_miss;
  }
}

__global__ void k_dense_ll(DenseBatch B, DenseState st, double* ll_out, int* status_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B.n_series) return;
  const int orig = B.order[b];
  double ll = 0;
  if (B.len_sorted[b] > 0)
    ll = st.zero[b] ? -DBL_MAX
                    : (log(st.p2[b]) - log(st.p1[b])) + (double)(st.e2[b] - st.e1[b]) * 0.693147180559945309417232121458;
  if (ll_out) ll_out[orig] = ll;
  if (status_out) status_out[orig] = B.len_sorted[b] > 0 ? st.bad[b] : 0;
}

// backward settle of slice t: rows whose last slice is t start with beta = 1, R = lambda_t;
// posterior of slice t = normalise(alpha_t * beta_t); h = 1 / (R_t . colsum) for the next GEMM.
__global__ void k_dense_bsettle(DenseBatch B, int t, int S, int SP, const double* colsum,
                                const double* lam_comb, const double* alpha, double* beta, double* R,
                                double* h, double* post, int post_stride, int post_off,
                                double* em_bt, double* em_h, double* em_r0, const double* phi0) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B.n_series || B.len_sorted[b] <= t) return;
  const long long r0 = B.row_off[B.order[b]];
  double* be = beta + (long long)b * SP;
  double* rr = R + (long long)b * SP;
  if (B.len_sorted[b] - 1 == t) {
    const double* l = lam_comb + (long long)B.cfg[r0 + t] * SP;
    for (int j = lane; j < SP; j += 32) { be[j] = j < S ? 1.0 : 0.0; rr[j] = l[j]; }
    __syncwarp();
  }
  const double* a = alpha + (r0 + t) * SP;
  double ps = 0, d = 0;
  for (int j = lane; j < SP; j += 32) { ps += a[j] * be[j]; d += rr[j] * colsum[j]; }
  ps = warp_sum_d(ps);
  d = warp_sum_d(d);
  const double pinv = rcp_or_one(ps);
  if (post) {
    double* prow = post + (r0 + t) * post_stride + post_off;
    for (int j = lane; j < S; j += 32) prow[j] = a[j] * be[j] * pinv;
  }
  if (lane == 0) h[b] = rcp_or_one(d);
  if (em_bt) {  // E-step: the carried beta_t, the scale of beta_{t-1}, and r_0 / (phi0 . r_0)
    for (int j = lane; j < SP; j += 32) em_bt[(r0 + t) * SP + j] = be[j];
    if (lane == 0) em_h[r0 + t] = rcp_or_one(d);
    if (t == 0) {
      double z = 0;
      for (int j = lane; j < SP; j += 32) z += phi0[j] * rr[j];
      const double zinv = rcp_or_one(warp_sum_d(z));
      for (int j = lane; j < SP; j += 32) em_r0[(long long)B.order[b] * SP + j] = rr[j] * zinv;
    }
  }
}

// ---- E-step statistics for large interfaces ---------------------------------------------
// N[k] = own_k . beta_k; one warp per data row
__global__ void k_dense_rownorm(const double* __restrict__ own, const double* __restrict__ bt, long long rows,
                                int SP, double* __restrict__ N) {
  const long long k = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= rows) return;
  double s = 0;
  for (int j = lane; j < SP; j += 32) s += own[k * SP + j] * bt[k * SP + j];
  s = warp_sum_d(s);
  if (lane == 0) N[k] = s;
}

// w[k] = h_k / N_{k-1}: weight of the pair (k-1, k); 0 on the rows that open a series
__global__ void k_dense_pairw(const double* __restrict__ N, const double* __restrict__ hv,
                              const unsigned char* __restrict__ first, long long rows, double* __restrict__ w) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= rows) return;
  const double n = k >= 1 ? N[k - 1] : 0.0;
  w[k] = (k >= 1 && !first[k] && n != 0) ? hv[k] / n : 0.0;
}

// posterior rows own_k beta_k / N_k added into evidence-indexed tables: CTA (x = 128-column
// slice, y = contiguous range of rows) keeps [n_comb][128] in shared memory, thread = column,
// rows in ascending order (deterministic); U rows are in flight per thread.
__global__ void __launch_bounds__(128) k_dense_leaf(const double* __restrict__ own, const double* __restrict__ bt,
                                                    const double* __restrict__ N, const int* __restrict__ cfg,
                                                    long long rows, int SP, int n_comb, double* __restrict__ part) {
  extern __shared__ double tabs[];
  for (int x = threadIdx.x; x < n_comb * 128; x += 128) tabs[x] = 0.0;
  __syncthreads();
  const int col = blockIdx.x * 128 + threadIdx.x;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long k_begin = blockIdx.y * per, k_end = k_begin + per < rows ? k_begin + per : rows;
  constexpr int U = 8;
  for (long long k0 = k_begin; k0 < k_end; k0 += U) {
    double o[U], b[U], n[U];
    int c[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long k = k0 + u;
      const bool ok = k < k_end;
      o[u] = ok ? own[k * SP + col] : 0.0;
      b[u] = ok ? bt[k * SP + col] : 0.0;
      n[u] = ok ? N[k] : 0.0;
      c[u] = ok ? cfg[k] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) tabs[c[u] * 128 + threadIdx.x] += n[u] != 0 ? (o[u] * b[u]) / n[u] : 0.0;
  }
  double* out = part + ((long long)blockIdx.y * n_comb) * SP + col;
  for (int c = 0; c < n_comb; c++) out[(long long)c * SP] = tabs[c * 128 + threadIdx.x];
}

// bt[k][j] <- beta_k[j] lambda_{c_k}[j] w[k]: the B operand of the count GEMM
__global__ void k_dense_bop(double* __restrict__ bt, const double* __restrict__ lam_comb,
                            const int* __restrict__ cfg, const double* __restrict__ w, long long rows, int SP) {
  const long long total = rows * (SP / 2);
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const long long k = x / (SP / 2);
    const int j = 2 * (int)(x - k * (SP / 2));
    double2 v = *reinterpret_cast<double2*>(bt + k * SP + j);
    const double2 l = *reinterpret_cast<const double2*>(lam_comb + (long long)cfg[k] * SP + j);
    const double wk = w[k];
    v.x *= l.x * wk;
    v.y *= l.y * wk;
    *reinterpret_cast<double2*>(bt + k * SP + j) = v;
  }
}

// G-part[z][i][j] = sum over the rows k of split z of A[k-1][i] * Bop[k][j]  (A = own rows):
// TN GEMM, 128 x 128 tile of G per CTA, 16 rows per step, cp.async double buffering, DMMA.
constexpr int CM = 128, CK = 16, CLD = CM + 4;
__global__ void __launch_bounds__(256) k_dense_counts(const double* __restrict__ own, const double* __restrict__ bop,
                                                      long long rows, int SP, double* __restrict__ part) {
  extern __shared__ double sm[];
  double* sA = sm;                    // [2][CK][CLD]
  double* sB = sm + 2 * CK * CLD;     // [2][CK][CLD]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  const int wm = w & 3, wn = w >> 2;  // warp tile: i 32*wm.., j 64*wn..
  const int i0 = blockIdx.y * CM, j0 = blockIdx.x * CM;
  long long chunk = (rows + gridDim.z - 1) / gridDim.z;
  chunk = (chunk + CK - 1) / CK * CK;
  const long long k_begin = (long long)blockIdx.z * chunk;
  const long long k_end = k_begin + chunk < rows ? k_begin + chunk : rows;
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto load_tiles = [&](int buf, long long k0) {
#pragma unroll
    for (int x = tid; x < CK * (CM / 2); x += 256) {
      const int r = x / (CM / 2), c = 2 * (x % (CM / 2));
      const long long k = k0 + r;
      const bool okb = k < k_end, oka = okb && k >= 1;
      cp_async16(sA + (buf * CK + r) * CLD + c, own + (oka ? (k - 1) * SP + i0 + c : 0), oka);
      cp_async16(sB + (buf * CK + r) * CLD + c, bop + (okb ? k * SP + j0 + c : 0), okb);
    }
    cp_async_commit();
  };
  if (k_begin < k_end) load_tiles(0, k_begin);
  int buf = 0;
  for (long long k0 = k_begin; k0 < k_end; k0 += CK, buf ^= 1) {
    if (k0 + CK < k_end) { load_tiles(buf ^ 1, k0 + CK); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* a_s = sA + buf * CK * CLD + 32 * wm;
    const double* b_s = sB + buf * CK * CLD + 64 * wn;
#pragma unroll
    for (int kk = 0; kk < CK / 4; kk++) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = a_s[(4 * kk + q) * CLD + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 8; j++) bf[j] = b_s[(4 * kk + q) * CLD + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
  double* out = part + (long long)blockIdx.z * SP * SP;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const long long o = (long long)(i0 + 32 * wm + 8 * i + g) * SP + j0 + 64 * wn + 8 * j + 2 * q;
      *reinterpret_cast<double2*>(out + o) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

__global__ void k_dense_mats(const double* base1, const int* ent_of, int S, int SP, double* A, double* AT) {
  const long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (x >= (long long)S * S) return;
  const int im = (int)(x / S), ip = (int)(x - (long long)im * S);
  const double v = base1[ent_of[x]];
  A[(long long)im * SP + ip] = v;     // forward:  k = previous state, n = current state
  AT[(long long)ip * SP + im] = v;    // backward: k = current state,  n = previous state
}

}  // namespace

int dense_refresh_mats(const ChainModel& cm, const double* d_base1_c0, cudaStream_t st) {
  const size_t n = (size_t)cm.SP * cm.SP;
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_Bf1, 0, n * sizeof(double), st));
  NIPGPU_CUDA(cudaMemsetAsync(cm.d_Bb1, 0, n * sizeof(double), st));
  const long long m = (long long)cm.S * cm.S;
  k_dense_mats<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_base1_c0, cm.d_ent_of, cm.S, cm.SP, cm.d_Bf1, cm.d_Bb1);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int dense_infer(const ChainModel& cm, ChainBatch& cb, const ChainPlan& plan, const ChainInferArgs& a,
                cudaStream_t st, const DenseEm* em) {
  const int S = cm.S, SP = cm.SP, n = a.n_series;
  if (n == 0) return NIPGPU_OK;
  // ---- per-batch buffers (sorted order) ----
  if (!cb.d_dense) {
    const size_t doubles = (size_t)n * (8 + 2 * (size_t)SP + 1 + 0) + 64;   // 7 state + h | beta | R(2 for ping-pong)
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_dense, ((size_t)n * (8 + 3 * (size_t)SP) + 64) * sizeof(double)));
    NIPGPU_CUDA(cudaMalloc((void**)&cb.d_dense_i, (size_t)n * 5 * sizeof(int) + 64));
    (void)doubles;
  }
  const size_t smem = (size_t)(2 * TM * LDA + 2 * TK * LDB) * sizeof(double);
  NIPGPU_CUDA(cudaFuncSetAttribute(k_dense_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NIPGPU_CUDA(cudaFuncSetAttribute(k_dense_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool filt = a.forward_only && a.d_post;
  double* fpost = filt ? a.d_post : nullptr;

  // the whole pass for the sorted positions [off, off + cnt) on stream s
  auto run_part = [&](int off, int cnt, cudaStream_t s) -> int {
    DenseState stt;
    double* p = cb.d_dense;
    stt.K = p + off; p += n; stt.Pp = p + off; p += n; stt.cp = p + off; p += n; stt.g = p + off; p += n;
    stt.p1 = p + off; p += n; stt.p2 = p + off; p += n;
    double* h = p + off; p += n; p += n;  // (spare)
    double* beta = p + (size_t)off * SP; p += (size_t)n * SP;
    double* R0 = p + (size_t)off * SP; p += (size_t)n * SP;
    double* R1b = p + (size_t)off * SP;
    int* ip = cb.d_dense_i;
    stt.e1 = ip + off; ip += n; stt.e2 = ip + off; ip += n; stt.zero = ip + off; ip += n; stt.bad = ip + off; ip += n;
    stt.noev = ip + off;
    DenseBatch B;
    B.n_series = cnt; B.order = cb.d_order + off; B.len_sorted = cb.d_len_sorted + off; B.row_off = a.d_row_off;
    B.cfg = cb.d_cfg;
    const int wgrid = (cnt + 7) / 8;  // 8 warps (sequences) per 256-thread block
    auto first_len = cb.len_sorted.begin() + off, last_len = first_len + cnt;
    auto n_longer = [&](int t) {      // sequences (a prefix of the sorted order) with length > t
      return (int)(std::lower_bound(first_len, last_len, t, std::greater<int>()) - first_len);
    };
    const int t_max = cnt > 0 ? *first_len : 0;   // longest sequence of this part

    k_dense_first<<<wgrid, 256, 0, s>>>(B, S, SP, cm.d_phi0, cb.d_comb, cm.m1_0, plan.c_miss, cb.d_alpha, stt);
    NIPGPU_LAUNCHED();
    for (int t = 1; t <= t_max; t++) {
      k_dense_fsettle<<<wgrid, 256, 0, s>>>(B, t, S, SP, cm.d_R1, plan.c_miss, cb.d_alpha, a.want_ll, fpost,
                                            a.post_stride, a.post_off, stt);
      NIPGPU_LAUNCHED();
      const int rows = n_longer(t);
      if (t < t_max && rows > 0) {
        dim3 grid(SP / TN, (rows + TM - 1) / TM);
        k_dense_gemm<0><<<grid, 256, smem, s>>>(B, t, rows, SP, cb.d_alpha, 1, cm.d_Bf1, cb.d_comb, stt.g,
                                                cb.d_alpha, nullptr);
        NIPGPU_LAUNCHED();
      }
    }
    k_dense_ll<<<(cnt + 255) / 256, 256, 0, s>>>(B, stt, a.want_ll ? a.d_ll : nullptr, a.d_status);
    NIPGPU_LAUNCHED();
    if (a.forward_only || (!a.d_post && !em)) return NIPGPU_OK;

    double* Rcur = R0;
    double* Rnext = R1b;
    for (int t = t_max - 1; t >= 0; t--) {
      k_dense_bsettle<<<wgrid, 256, 0, s>>>(B, t, S, SP, cm.d_colsum, cb.d_comb, cb.d_alpha, beta, Rcur, h, a.d_post,
                                            a.post_stride, a.post_off, em ? em->bt : nullptr,
                                            em ? em->hvec : nullptr, em ? em->r0 : nullptr, cm.d_phi0);
      NIPGPU_LAUNCHED();
      const int rows = n_longer(t);
      if (t >= 1 && rows > 0) {
        dim3 grid(SP / TN, (rows + TM - 1) / TM);
        k_dense_gemm<1><<<grid, 256, smem, s>>>(B, t, rows, SP, Rcur, 0, cm.d_Bb1, cb.d_comb, h, beta, Rnext);
        NIPGPU_LAUNCHED();
        std::swap(Rcur, Rnext);
      }
    }
    return NIPGPU_OK;
  };

  // Parts of the batch on their own streams: one part's settle kernels and the ragged last wave
  // of its GEMMs (256 CTAs on 148 SMs for the whole C4 batch) run under the other parts' GEMMs.
  // Sorted by length: interleaving would balance ragged sets better, contiguous parts keep the
  // views plain pointer offsets.
  int parts = n >= 512 ? 2 : 1;   // C4: 72 ms in one part, 62 ms in two, 67 / 66 ms in three / four
  if (const char* e = getenv("NIPGPU_DENSE_PARTS")) parts = std::max(1, std::min(4, atoi(e)));
  if (parts == 1) return run_part(0, n, st);
  if (!cb.dense_fork) NIPGPU_CUDA(cudaEventCreateWithFlags(&cb.dense_fork, cudaEventDisableTiming));
  const int per = ((n + parts - 1) / parts + 7) / 8 * 8;
  NIPGPU_CUDA(cudaEventRecord(cb.dense_fork, st));
  for (int k = 1; k < parts; k++) {
    const int off = k * per, cnt = std::min(per, n - off);
    if (cnt <= 0) break;
    if (!cb.dense_stream[k - 1]) {
      NIPGPU_CUDA(cudaStreamCreateWithFlags(&cb.dense_stream[k - 1], cudaStreamNonBlocking));
      NIPGPU_CUDA(cudaEventCreateWithFlags(&cb.dense_join[k - 1], cudaEventDisableTiming));
    }
    NIPGPU_CUDA(cudaStreamWaitEvent(cb.dense_stream[k - 1], cb.dense_fork, 0));
    if (int e = run_part(off, cnt, cb.dense_stream[k - 1])) return e;
    NIPGPU_CUDA(cudaEventRecord(cb.dense_join[k - 1], cb.dense_stream[k - 1]));
  }
  if (int e = run_part(0, std::min(per, n), st)) return e;
  for (int k = 1; k < parts; k++)
    if (k * per < n) NIPGPU_CUDA(cudaStreamWaitEvent(st, cb.dense_join[k - 1], 0));
  return NIPGPU_OK;
}

int dense_stats(const ChainModel& cm, const ChainBatch& cb, const ChainPlan& plan, const DenseEm& em,
                const unsigned char* first, long long rows, double* work, int partsG, double* partG,
                int partsC, double* partC, cudaStream_t st) {
  const int SP = cm.SP;
  double* N = work;
  double* w = work + rows;
  if (rows <= 0) {
    NIPGPU_CUDA(cudaMemsetAsync(partG, 0, (size_t)partsG * SP * SP * sizeof(double), st));
    NIPGPU_CUDA(cudaMemsetAsync(partC, 0, (size_t)partsC * plan.n_comb * SP * sizeof(double), st));
    return NIPGPU_OK;
  }
  k_dense_rownorm<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(cb.d_alpha, em.bt, rows, SP, N);
  NIPGPU_LAUNCHED();
  k_dense_pairw<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(N, em.hvec, first, rows, w);
  NIPGPU_LAUNCHED();
  {
    const size_t smem = (size_t)plan.n_comb * 128 * sizeof(double);
    if (smem > 200 * 1024) return NIPGPU_EUNSUPPORTED;
    NIPGPU_CUDA(cudaFuncSetAttribute(k_dense_leaf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dense_leaf<<<dim3(SP / 128, partsC), 128, smem, st>>>(cb.d_alpha, em.bt, N, cb.d_cfg, rows, SP, plan.n_comb, partC);
    NIPGPU_LAUNCHED();
  }
  k_dense_bop<<<148 * 8, 256, 0, st>>>(em.bt, cb.d_comb, cb.d_cfg, w, rows, SP);
  NIPGPU_LAUNCHED();
  {
    const size_t smem = (size_t)4 * CK * CLD * sizeof(double);
    NIPGPU_CUDA(cudaFuncSetAttribute(k_dense_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dense_counts<<<dim3(SP / CM, SP / CM, partsG), 256, smem, st>>>(cb.d_alpha, em.bt, rows, SP, partG);
    NIPGPU_LAUNCHED();
  }
  return NIPGPU_OK;
}

}  // namespace nipgpu
