// chain_pair.cuh — the forward / backward recursions of chain.cu with every group of 8 sequences
// split over TWO warps that sit on the same scheduler (included by chain.cu, inside its
// anonymous namespace; NT = 4 or 8 state tiles).
//
// Why: the recurrence is sequential in t, and with one warp per scheduler everything that is
// not a tensor instruction — B-fragment loads waiting for the read barrier of the DMMA whose
// registers they re-use, the side work between the sweeps, the hand-over from one sweep to the
// next — is exposed: 23.8 clocks per DMMA in the one-warp kernels, 18.9 for the bare sweep
// (nipgpu_probe_sweep), against a pipe that takes one every 16.  Two warps on one scheduler
// fill each other's bubbles (16.5 clocks per DMMA for two independent sweeps, same probe).  A
// batch of 4096 sequences has only one 8-sequence group per scheduler, so the second warp has
// to come from splitting the group's own work:
//
//   warp h (0 / 1) of a pair owns the state tiles [h NH, (h+1) NH), NH = NT/2: it computes that
//   half of every result vector — NH accumulator chains per k-step — and needs the FULL vector
//   as the A operand of the next sweep.  Its own half is in registers (accumulator layout = A
//   operand layout, as in the one-warp kernels); the partner's half comes through shared memory,
//   lane to same lane (the value a lane needs for k-step (n, e) of the partner's half is the one
//   the partner's same lane holds in tile n, element e).  Each sweep starts with the k-steps of
//   the warp's own half; the halves are written in the first side slots of the sweep, a named
//   barrier of the pair's 64 threads follows a quarter into the sweep, and the partner's half is
//   read well before the second half of the k loop wants it: the exchange costs no latency.
//   Scalars (masses, scales, likelihood terms) are computed by both warps from the same
//   numbers in the same order, so they agree bit for bit and need no exchange: sums over the 64
//   states are formed as (partial over one half) + (partial over the other), both partials
//   computed by the half's owner order in both warps, and a + b == b + a.
//
// Scalar FP64 instructions run on the SAME pipe as DMMA.8x8x4 on this chip (a warp-wide DFMA
// takes ~2.6 of its clocks, nipgpu_probe_dmma_dfma), so every DMUL / DADD / DSETP of the side
// work is paid in tensor time; the one-warp kernels spend ~230 of them per slice (lagged
// scaling, two multiplies per element and a reciprocal per slice).  These kernels keep only
// what the slice needs:
//   * the vectors are NOT normalised per slice.  When the largest binary exponent of a vector
//     leaves [-60, 60] (integer compares on the high words, no FP64 instruction) the next one
//     is scaled by the exact power of two that brings it back — a warp-uniform, rarely taken
//     branch; powers of two change no mantissa, so the results do not depend on when it fires;
//   * the log-likelihood sum_t log m2_t - log m1_t telescopes: with c_t = sum of the vector
//     and d_t = vector . R1,  m2_t / m1_t = c_t / (f_t d_{t-1})  (f_t the scale applied at t),
//     accumulated as mantissa products with binary exponents (LogAcc);
//   * posterior of slice t = normalise(alpha_t * beta_t), its normaliser from the two halves.
#pragma once

// Where the side items of a sweep sit (slots of tensor instructions): part A in [0, A_END), part
// B in [B_START, B_END), the partner's half fetched after slot MID = NS/4.  The placement changes
// the kernels' time by up to 15 % (the scalar FP64 items share the tensor pipe and each waits
// behind the DMMAs in flight); the values below are the best of a sweep on the C2 shape
// (tools/ab_variants.py, profiles/r02_pair_item_placement.txt).  NS = 64 for 8 state tiles.
#ifndef NIPGPU_PAIR_A_END
#define NIPGPU_PAIR_A_END (MID / 4 > 0 ? MID / 4 : 1)
#endif
#ifndef NIPGPU_PAIR_B_START
#define NIPGPU_PAIR_B_START (MID + 2)
#endif
#ifndef NIPGPU_PAIR_B_END
#define NIPGPU_PAIR_B_END (MID + 2 + NS / 8)
#endif
#ifndef NIPGPU_PAIR_FWD_A_END
#define NIPGPU_PAIR_FWD_A_END NIPGPU_PAIR_A_END
#define NIPGPU_PAIR_FWD_B_START NIPGPU_PAIR_B_START
#define NIPGPU_PAIR_FWD_B_END NIPGPU_PAIR_B_END
#endif
#ifndef NIPGPU_PAIR_BWD_A_END
#define NIPGPU_PAIR_BWD_A_END NIPGPU_PAIR_A_END
#define NIPGPU_PAIR_BWD_B_START NIPGPU_PAIR_B_START
#define NIPGPU_PAIR_BWD_B_END NIPGPU_PAIR_B_END
#endif

// W = 2: the two warps of a team sit on ONE scheduler (warps p and p + 4 of an 8-warp CTA, four
//        teams per CTA); they fill each other's bubbles on that scheduler's pipe.
// W = 4: the four warps of a team sit on the FOUR schedulers of an SM (a CTA is one team); a
//        group's slice then takes a quarter of the tensor time, which is what a small batch
//        needs (one group per SM still uses all four pipes), and a large batch puts four
//        CTAs on an SM, i.e. four warps of four independent teams on every scheduler.
template <int NT, int W>
struct TeamGeom {
  static constexpr int SP = 8 * NT, NH = NT / W;   // NH state tiles per warp
  static constexpr int TEAMS = W == 2 ? 4 : 1;     // teams per CTA
  static constexpr int THREADS = 32 * W * TEAMS;
  static constexpr int NS = 2 * NT * NH;           // tensor instructions (= side slots) per warp and sweep
  static constexpr int OWN = 2 * NH * NH;          // of which on the warp's own part of the vector
  static constexpr int MID = OWN / 2;              // slot after which the other parts are fetched
  static constexpr int REC = NH + 1;               // double2 per lane in the exchange record: part + two partial sums
  // exchange area in double2: [team][parity 2][part W][REC][32 lanes]
  static constexpr int XCH = TEAMS * 2 * W * REC * 32;
  static size_t smem_bytes() { return sizeof(double) * (SP * SP + SP) + sizeof(double2) * XCH; }
};

template <int W>
__device__ __forceinline__ void team_barrier(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(32 * W) : "memory");
}

// acc[n] (tile h NH + n) = sum over all 2 NT k-steps; the k-steps of the warp's own part first,
// then the other parts in the order h+1, h+2, ... (mod W).
//   fp[i]     B fragments of this warp's n-tiles for the k-steps of part i (double2 pointers,
//             lane already added)
//   side(slot) after every tensor instruction; side(MID) must have filled `theirs`
template <int NT, int W, class Side>
__device__ __forceinline__ void team_sweep(double (&acc)[NT / W][2], const double (&mine)[NT / W][2],
                                           const double (&theirs)[W - 1][NT / W][2],
                                           const double2* const (&fp)[W], Side side) {
  constexpr int NH = NT / W, H2 = NH / 2, KP = 2 * NH;   // H2 double2 loads bring the B fragments of a k-step
  double b[2][NH];
  auto fetch = [&](auto ksc) {
    constexpr int ks = decltype(ksc)::value, buf = ks & 1;
    constexpr int part = ks / KP, kl = ks % KP;          // k-step inside its part
    const double2* p = fp[part] + ((kl * (NT / 2)) << 5);
    static_for<0, H2>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      const double2 v = p[m << 5];
      b[buf][2 * m] = v.x;
      b[buf][2 * m + 1] = v.y;
    });
  };
  fetch(std::integral_constant<int, 0>{});
  static_for<0, 2 * NT>([&](auto ksc) {
    constexpr int ks = decltype(ksc)::value;
    if constexpr (ks + 1 < 2 * NT) fetch(std::integral_constant<int, ks + 1>{});
    constexpr int part = ks / KP, kl = ks % KP;
    double av;
    if constexpr (part == 0) av = mine[kl >> 1][kl & 1];
    else av = theirs[part - 1][kl >> 1][kl & 1];
    static_for<0, NH>([&](auto nc) {
      constexpr int n = decltype(nc)::value;
      if constexpr (ks == 0) dmma_init(acc[n][0], acc[n][1], av, b[0][n]);
      else dmma(acc[n][0], acc[n][1], av, b[ks & 1][n]);
      side(std::integral_constant<int, ks * NH + n>{});
    });
  });
}

// items w of a list of W spread over the slots [LO, HI) of a sweep
template <int LO, int HI, int W, int SLOT, class Item>
__device__ __forceinline__ void run_items_in(Item& item) {
  if constexpr (SLOT >= LO && SLOT < HI)
    static_for<0, W>([&](auto wc) {
      if constexpr ((decltype(wc)::value * (HI - LO)) / W == SLOT - LO) item(wc);
    });
}

// ---------------------------------------------------------------- forward ---
// largest biased binary exponent among the (non-negative) values, as an integer
__device__ __forceinline__ int hi_word(double x) { return __double2hiint(x); }
// f = 2^-(e - 1023) for a biased exponent e far from 1023, else 1 (e == 0: a zero vector stays)
__device__ __forceinline__ double pow2_rescale(int maxhi, bool& need) {
  const int e = (maxhi >> 20) & 0x7ff;
  need = e != 0 && (e < 1023 - 60 || e > 1023 + 60);
  return need ? __hiloint2double((2046 - e) << 20, 0) : 1.0;
}

// 4-byte read-only load that stays where it is written (see ldg_pinned in chain.cu)
__device__ __forceinline__ int ldg_pinned_int(const int* p, bool on) {
  int v = 0;
  if (on) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// exponent the rescale adds to the vector's (0 when none is needed)
__device__ __forceinline__ int pow2_rescale_exp(int maxhi) {
  const int e = (maxhi >> 20) & 0x7ff;
  return (e != 0 && (e < 1023 - 60 || e > 1023 + 60)) ? 1023 - e : 0;
}
// 2^k as a double, k clamped to the normal range
__device__ __forceinline__ double pow2_double(int k) {
  k = k < -1000 ? -1000 : (k > 1000 ? 1000 : k);
  return __hiloint2double((1023 + k) << 20, 0);
}

// m <= 0 / m == 0 for a finite double, on the integer pipe
__device__ __forceinline__ bool le_zero(double m) {
  const int hi = __double2hiint(m), lo = __double2loint(m);
  return hi < 0 || (hi | lo) == 0;
}
__device__ __forceinline__ bool is_zero(double m) {
  return ((__double2hiint(m) & 0x7fffffff) | __double2loint(m)) == 0;
}

// LogAcc::add with the sign tests done on the integer pipe
__device__ __forceinline__ void logacc_add(LogAcc& L, double m1, double m2, bool on) {
  const bool p1ok = !le_zero(m1), p2ok = !le_zero(m2);
  const bool both = on && p1ok && p2ok;
  L.p1 *= both ? m1 : 1.0;
  L.p2 *= both ? m2 : 1.0;
  LogAcc::renorm(L.p1, L.e1);
  LogAcc::renorm(L.p2, L.e2);
  if (on && is_zero(m2)) L.zero = 1;
  const bool pos = L.e2 > L.e1 || (L.e2 == L.e1 && L.p2 > L.p1);  // running log-likelihood > 0
  if (on && (!p1ok || !p2ok || (pos && !L.zero))) L.bad = 1;
}

template <int NT, int W, bool FILT, bool WLL>
__global__ void __launch_bounds__(32 * W * (W == 2 ? 4 : 1), W == 2 ? 1 : 4) k_chain_forward_team(ChainDev C, ChainBatchDev B,
                                                               double* __restrict__ alpha,
                                                               double* __restrict__ post, int post_stride,
                                                               int post_off, double* ll_out, int* status_out) {
  using G = TeamGeom<NT, W>;
  constexpr int SP = G::SP, NH = G::NH, NS = G::NS, MID = G::MID, REC = G::REC, TEAMS = G::TEAMS;
  constexpr bool SUMS = WLL || FILT;
  extern __shared__ double sB[];
  double* s_r1 = sB + SP * SP;
  double2* s_x = reinterpret_cast<double2*>(s_r1 + SP);
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) sB[i] = C.Bf1[i];
  for (int i = threadIdx.x; i < SP; i += blockDim.x) s_r1[i] = C.R1[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int wib = threadIdx.x >> 5, pair = wib % TEAMS, h = wib / TEAMS;   // W = 2: warps p and p + 4 share scheduler p
  const int bp = (blockIdx.x * TEAMS + pair) * 8 + g;
  const bool valid = bp < B.n_series;
  const int T = valid ? B.len_sorted[bp] : 0;
  const int Tw = __shfl_sync(0xffffffffu, T, 0);  // sorted by length: row 0 is the longest
  const int orig = valid ? B.order[bp] : 0;
  const long long row0 = valid ? B.row_off[orig] : 0;
  const int* cfg = B.cfg + row0;
  const int t0 = h * NH;                          // first state tile of the warp's own part
  const double2* f2 = reinterpret_cast<const double2*>(sB) + lane + ((h * (NH / 2)) << 5);
  const double2* fp[W];                           // B fragments for the k-steps of part h, h+1, ... (mod W)
#pragma unroll
  for (int i = 0; i < W; i++) fp[i] = f2 + ((2 * (((h + i) % W) * NH) * (NT / 2)) << 5);
  const double2* r1v = reinterpret_cast<const double2*>(s_r1);
  // exchange records: xw(parity) is written by this warp, xr(parity, part) by the warp owning `part`
  double2* xbase = s_x + (size_t)pair * (2 * W * REC * 32) + lane;
  auto xw = [&](int par) { return xbase + (par * W + h) * (REC * 32); };
  auto xr = [&](int par, int part) { return xbase + (par * W + part) * (REC * 32); };

  double mine[NH][2], theirs[W - 1][NH][2], acc[NH][2], lam[NH][2];
  LogAcc L;
  auto load_lam = [&](int c, bool on) {   // own half of the evidence row
    const double2* p = reinterpret_cast<const double2*>(C.lam_comb + (long long)c * SP);
#pragma unroll
    for (int n = 0; n < NH; n++) {
      const double2 v = ldg_pinned(p + 4 * (t0 + n) + q, on);
      lam[n][0] = v.x;
      lam[n][1] = v.y;
    }
  };
  // slice 0: the vector is phi0 * lambda_0 as it is; its mass is m2_0, m1_0 is a model constant
  int c_cur = T > 0 ? cfg[0] : 0;
  int c_next = T > 1 ? cfg[1] : 0;
  load_lam(c_cur, T > 0);
#pragma unroll
  for (int n = 0; n < NH; n++) {
    mine[n][0] = C.phi0[8 * (t0 + n) + 2 * q] * lam[n][0];
    mine[n][1] = C.phi0[8 * (t0 + n) + 2 * q + 1] * lam[n][1];
  }
  double m1n = C.m1_0;     // m1 of the slice being settled, times the common factor of its m2
  int ea = 0;              // binary exponent the vector being settled carries (sum of the rescales)
  bool noev_p = c_cur == C.c_miss, on_p = T > 0;

  // Side work of one sweep.  Part A (first quarter of the slots) touches only the own half of the
  // vector being settled (slice s = t-1, the sweep's A operand); part B runs after the partner's
  // half has arrived.
  constexpr int EH = 2 * NH;
  constexpr int A_SUM = 0;                    // EH items : partial sum of the own half
  constexpr int A_DOT = A_SUM + EH;           // NH items : own half . R1
  constexpr int A_X = A_DOT + NH;             // NH + 1   : own half and the two partials -> exchange record
  constexpr int A_ST = A_X + NH + 1;          // NH items : alpha row, own half
  constexpr int A_EXP = A_ST + NH;            // 1 item   : largest exponent of the own half
  constexpr int WA = A_EXP + 1;
  constexpr int B_EXP = 0;                    // 3 items  : largest exponent of the vector, the power of two
  constexpr int B_RED = B_EXP + 3;            // 3 items  : c = sum of the vector
  constexpr int B_DRED = B_RED + 3;           // 3 items  : d = vector . R1
  constexpr int B_LL = B_DRED + 3;            // 1 item   : likelihood bookkeeping
  constexpr int B_FILT = B_LL + 1;            // NH items : filtered marginal, own half
  constexpr int B_F = B_FILT + NH;            // 1 item   : rescale (rare)
  constexpr int WB = B_F + 1;
  double pa[2], da[2], part_c = 0, part_d = 0, cs = 0, ds = 0, cinv = 1.0, fscale = 1.0;
  int s_slice = 0, par = 0, mx = 0, fk = 0;
  bool need = false, scaled = false;
  auto item_a = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w >= A_SUM && w < A_DOT) {
      if constexpr (SUMS) {
        constexpr int i = w - A_SUM;
        if constexpr (i < 2) pa[i] = mine[i >> 1][i & 1];
        else pa[i & 1] += mine[i >> 1][i & 1];
      }
    } else if constexpr (w >= A_DOT && w < A_X) {
      if constexpr (WLL) {
        constexpr int n = w - A_DOT;
        const double2 v = r1v[4 * (t0 + n) + q];
        if constexpr (n == 0) { da[0] = mine[n][0] * v.x; da[1] = mine[n][1] * v.y; }
        else { da[0] = fma(mine[n][0], v.x, da[0]); da[1] = fma(mine[n][1], v.y, da[1]); }
      }
    } else if constexpr (w >= A_X && w < A_X + NH) {
      constexpr int n = w - A_X;
      xw(par)[n << 5] = make_double2(mine[n][0], mine[n][1]);
    } else if constexpr (w == A_X + NH) {
      if constexpr (SUMS) part_c = pa[0] + pa[1];
      if constexpr (WLL) part_d = da[0] + da[1];
      if constexpr (SUMS) xw(par)[NH << 5] = make_double2(part_c, part_d);
    } else if constexpr (w >= A_ST && w < A_EXP) {
      constexpr int n = w - A_ST;
      if (on_p) reinterpret_cast<double2*>(alpha + (row0 + s_slice) * SP)[4 * (t0 + n) + q] = make_double2(mine[n][0], mine[n][1]);
      if constexpr (n == 0) {
        if (on_p && q == 0 && h == 0) B.fexp[row0 + s_slice] = ea;
      }
    } else if constexpr (w == A_EXP) {
      mx = 0;
#pragma unroll
      for (int n = 0; n < NH; n++) mx = max(mx, max(hi_word(mine[n][0]), hi_word(mine[n][1])));
    }
  };
  auto mid = [&]() {   // the other parts of the vector being settled
    team_barrier<W>(pair);
#pragma unroll
    for (int i = 1; i < W; i++) {
      const double2* rec = xr(par, (h + i) % W);
#pragma unroll
      for (int n = 0; n < NH; n++) {
        const double2 v = rec[n << 5];
        theirs[i - 1][n][0] = v.x;
        theirs[i - 1][n][1] = v.y;
      }
    }
  };
  auto item_b = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w == B_EXP) {
#pragma unroll
      for (int i = 0; i < W - 1; i++)
#pragma unroll
        for (int n = 0; n < NH; n++) mx = max(mx, max(hi_word(theirs[i][n][0]), hi_word(theirs[i][n][1])));
    } else if constexpr (w == B_EXP + 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    } else if constexpr (w == B_EXP + 2) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      fscale = pow2_rescale(mx, need);
      fk = pow2_rescale_exp(mx);
    } else if constexpr (w == B_RED) {
      if constexpr (SUMS) {   // the parts' partial sums in the order of the parts: the same sum in every warp
        double c = 0, d = 0;
#pragma unroll
        for (int part = 0; part < W; part++) {
          const double2 v = xr(par, part)[NH << 5];
          c = part == 0 ? v.x : c + v.x;
          d = part == 0 ? v.y : d + v.y;
        }
        cs = c;
        ds = d;
      }
    } else if constexpr (w == B_RED + 1) {
      if constexpr (SUMS) cs += __shfl_xor_sync(0xffffffffu, cs, 1);
    } else if constexpr (w == B_RED + 2) {
      if constexpr (SUMS) cs += __shfl_xor_sync(0xffffffffu, cs, 2);
    } else if constexpr (w == B_DRED) {
    } else if constexpr (w == B_DRED + 1) {
      if constexpr (WLL) ds += __shfl_xor_sync(0xffffffffu, ds, 1);
    } else if constexpr (w == B_DRED + 2) {
      if constexpr (WLL) ds += __shfl_xor_sync(0xffffffffu, ds, 2);
    } else if constexpr (w == B_LL) {
      if constexpr (WLL) {
        // slice s: m1 = m1n, m2 = c_s, both times the same positive factor (src/nip.c:1458-1474);
        // without evidence m2 := m1 (deviations)
        logacc_add(L, m1n, noev_p ? m1n : cs, on_p);
        m1n = ds * fscale;   // for slice s+1: d_s times the scale its vector is about to receive
      }
      if constexpr (FILT) cinv = safe_rcp(cs);
    } else if constexpr (w >= B_FILT && w < B_F) {
      if constexpr (FILT) {  // filtering: the forward marginal of I_s is the vector over its sum
        constexpr int n = w - B_FILT;
        double* prow = post + (row0 + s_slice) * post_stride + post_off;
        const int col = 8 * (t0 + n) + 2 * q;
        if (on_p && col < C.S) prow[col] = mine[n][0] * cinv;
        if (on_p && col + 1 < C.S) prow[col + 1] = mine[n][1] * cinv;
      }
    } else if constexpr (w == B_F) {
      scaled = __any_sync(0xffffffffu, need);   // a vote, no branch inside the sweep
    }
  };
  auto side = [&](auto sc) {
    constexpr int s = decltype(sc)::value;
    run_items_in<0, NIPGPU_PAIR_FWD_A_END, WA, s>(item_a);
    if constexpr (s == MID) mid();
    run_items_in<NIPGPU_PAIR_FWD_B_START, NIPGPU_PAIR_FWD_B_END, WB, s>(item_b);
  };

  c_cur = c_next;
  load_lam(c_cur, 1 < T);
  if (2 < T) c_next = __ldg(cfg + 2);
  for (int t = 1; t < Tw; t++) {
    const bool on = t < T;
    s_slice = t - 1;
    par = (t - 1) & 1;
    team_sweep<NT, W>(acc, mine, theirs, fp, side);
    noev_p = c_cur == C.c_miss; on_p = on;
#pragma unroll
    for (int n = 0; n < NH; n++) { mine[n][0] = acc[n][0] * lam[n][0]; mine[n][1] = acc[n][1] * lam[n][1]; }
    if (scaled) {   // warp-uniform and rare: an exact power of two per sequence (1 where not needed)
#pragma unroll
      for (int n = 0; n < NH; n++) { mine[n][0] *= fscale; mine[n][1] *= fscale; }
    }
    ea += fk;       // exponent of the vector of slice t
    c_cur = c_next;
    load_lam(c_cur, t + 1 < T);     // evidence row of slice t+1
    if (t + 2 < T) c_next = __ldg(cfg + t + 2);
  }
  if (Tw >= 1) {   // settle the last slice: the same items, back to back
    s_slice = Tw - 1;
    par = (Tw - 1) & 1;
    static_for<0, WA>(item_a);
    mid();
    static_for<0, WB>(item_b);
  }
  if (valid && q == 0 && h == 0) {
    if (ll_out) ll_out[orig] = (WLL && T > 0) ? L.value() : 0.0;
    if (status_out) status_out[orig] = L.bad;
  }
}

// --------------------------------------------------------------- backward ---
// beta_{t-1} = f_t A (lambda_t * beta_t) with f_t an exact power of two (1 almost always);
// posterior of slice t = alpha_t * beta_t / (alpha_t . beta_t), the normaliser from the scale
// bookkeeping (no sum, no reciprocal per slice).  Exchange record: the own half of
// r_t = lambda_t * beta_t (the sweep's A operand).
// EM variant: stores the carried beta_t (rt[t]) and f_t (hvec[t]) instead of posteriors, and
// r_0 / (phi0 . r_0) per series; see k_chain_backward and k_chain_stats.
template <int NT, int W, bool VEC, bool EM>
__global__ void __launch_bounds__(32 * W * (W == 2 ? 4 : 1), W == 2 ? 1 : 4) k_chain_backward_team(ChainDev C, ChainBatchDev B,
                                                                const double* __restrict__ alpha,
                                                                double* __restrict__ post, int post_stride,
                                                                int post_off, double* __restrict__ rt,
                                                                double* __restrict__ r0,
                                                                double* __restrict__ hvec) {
  using G = TeamGeom<NT, W>;
  constexpr int SP = G::SP, NH = G::NH, NS = G::NS, MID = G::MID, REC = G::REC, TEAMS = G::TEAMS;
  extern __shared__ double sB[];
  double2* s_x = reinterpret_cast<double2*>(sB + SP * SP + SP);
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) sB[i] = C.Bb1[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int wib = threadIdx.x >> 5, pair = wib % TEAMS, h = wib / TEAMS;
  const int bp = (blockIdx.x * TEAMS + pair) * 8 + g;
  const bool valid = bp < B.n_series;
  const int T = valid ? B.len_sorted[bp] : 0;
  const int Tw = __shfl_sync(0xffffffffu, T, 0);
  const int orig = valid ? B.order[bp] : 0;
  const long long row0 = valid ? B.row_off[orig] : 0;
  const int* cfg = B.cfg + row0;
  const int t0 = h * NH;
  const double2* f2 = reinterpret_cast<const double2*>(sB) + lane + ((h * (NH / 2)) << 5);
  const double2* fp[W];
#pragma unroll
  for (int i = 0; i < W; i++) fp[i] = f2 + ((2 * (((h + i) % W) * NH) * (NT / 2)) << 5);
  double2* xbase = s_x + (size_t)pair * (2 * W * REC * 32) + lane;
  auto xw = [&](int par) { return xbase + (par * W + h) * (REC * 32); };
  auto xr = [&](int par, int part) { return xbase + (par * W + part) * (REC * 32); };

  // own halves: beta_t, r_t = lambda_t * beta_t (A operand), alpha_t, evidence row of slice t-1
  double beta[NH][2], r[NH][2], rth[W - 1][NH][2], u[NH][2], lam[NH][2], a[NH][2], an[NH][2];
  auto load_half = [&](const double* base, bool on, double (&dst)[NH][2]) {
    const double2* p = reinterpret_cast<const double2*>(base);
#pragma unroll
    for (int n = 0; n < NH; n++) {
      const double2 v = ldg_pinned(p + 4 * (t0 + n) + q, on);
      dst[n][0] = v.x;
      dst[n][1] = v.y;
    }
  };
  // prologue: the longest rows start at slice Tw-1 with beta = 1, r = lambda
  const bool has_last = Tw >= 1 && Tw - 1 < T;
  load_half(C.lam_comb + (long long)(has_last ? cfg[Tw - 1] : 0) * SP, has_last, r);
  if (!EM) load_half(alpha + (row0 + Tw - 1) * SP, has_last, a);
#pragma unroll
  for (int n = 0; n < NH; n++) beta[n][0] = beta[n][1] = 1.0;
  int c_pre = (Tw >= 2 && Tw - 2 < T) ? cfg[Tw - 2] : 0;

  // The normaliser of the posterior needs no sum: forward row . beta row = Z 2^(Fa_t + Fb_t) for
  // every t, with Z = sum of the last forward row / 2^(its exponent) (k_chain_final) and Fa / Fb
  // the exponents the two recursions have put on their vectors.
  const double rbase = safe_rcp(valid ? B.zc[orig] : 0.0);
  const int zf = valid ? B.zf[orig] : 0;
  int eb = 0;                                 // exponent carried by beta_t
  int fa_cur = has_last ? B.fexp[row0 + Tw - 1] : 0, fa_nxt = 0, fa_pre = 0;   // exponents of the forward rows t, t-1, t-2

  constexpr int A_MUL = 0;                    // NH items : a = alpha_t * beta_t (own half)
  constexpr int A_X = A_MUL + NH;             // NH items : exchange record (r half)
  constexpr int A_EXP = A_X + NH;             // 1 item   : largest exponent of the own half of r
  constexpr int A_RN = A_EXP + 1;             // 1 item   : 1 / (alpha_t . beta_t)
  constexpr int WA = A_RN + 1;
  constexpr int B_EXP = 0;                    // 3 items  : largest exponent of r, the power of two
  constexpr int B_ST = B_EXP + 3;             // NH items : posterior of slice t / the carried beta (EM)
  constexpr int B_LD = B_ST + NH;             // NH items : alpha_{t-1} (prefetched) -> a
  constexpr int B_F = B_LD + NH;              // 1 item   : rescale vote, scalars of the E-step
  constexpr int WB = B_F + 1;
  double pinv = 1.0, fscale = 1.0;
  int t_cur = 0, par = 0, mx = 0, fk = 0;
  bool on = false, first_next = false, need = false, scaled = false;
  auto item_a = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w >= A_MUL && w < A_X) {
      if constexpr (!EM) {
        constexpr int n = w - A_MUL;
        a[n][0] *= beta[n][0];
        a[n][1] *= beta[n][1];
      }
    } else if constexpr (w >= A_X && w < A_EXP) {
      constexpr int n = w - A_X;
      xw(par)[n << 5] = make_double2(r[n][0], r[n][1]);
    } else if constexpr (w == A_EXP) {
      mx = 0;
#pragma unroll
      for (int n = 0; n < NH; n++) mx = max(mx, max(hi_word(r[n][0]), hi_word(r[n][1])));
    } else if constexpr (w == A_RN) {
      pinv = rbase * pow2_double(zf - fa_cur - eb);
    }
  };
  auto mid = [&]() {
    team_barrier<W>(pair);
#pragma unroll
    for (int i = 1; i < W; i++) {
      const double2* rec = xr(par, (h + i) % W);
#pragma unroll
      for (int n = 0; n < NH; n++) {
        const double2 v = rec[n << 5];
        rth[i - 1][n][0] = v.x;
        rth[i - 1][n][1] = v.y;
      }
    }
  };
  auto item_b = [&](auto wc) {
    constexpr int w = decltype(wc)::value;
    if constexpr (w == B_EXP) {
#pragma unroll
      for (int i = 0; i < W - 1; i++)
#pragma unroll
        for (int n = 0; n < NH; n++) mx = max(mx, max(hi_word(rth[i][n][0]), hi_word(rth[i][n][1])));
    } else if constexpr (w == B_EXP + 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    } else if constexpr (w == B_EXP + 2) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      fscale = pow2_rescale(mx, need);
      fk = pow2_rescale_exp(mx);
      if (first_next) { fscale = 1.0; need = false; fk = 0; }   // slice t-1 restarts this row with beta = 1
    } else if constexpr (w >= B_ST && w < B_LD) {
      constexpr int n = w - B_ST;
      if constexpr (EM) {  // E-step: the carried beta_t itself
        if (on) reinterpret_cast<double2*>(rt + (row0 + t_cur) * SP)[4 * (t0 + n) + q] = make_double2(beta[n][0], beta[n][1]);
      } else {
        double* prow = post + (row0 + t_cur) * post_stride + post_off;
        if (VEC) {
          if (on) reinterpret_cast<double2*>(prow)[4 * (t0 + n) + q] = make_double2(a[n][0] * pinv, a[n][1] * pinv);
        } else {
          const int col = 8 * (t0 + n) + 2 * q;
          if (on && col < C.S) prow[col] = a[n][0] * pinv;
          if (on && col + 1 < C.S) prow[col + 1] = a[n][1] * pinv;
        }
      }
    } else if constexpr (w >= B_LD && w < B_F) {
      if constexpr (!EM) {
        constexpr int n = w - B_LD;
        a[n][0] = an[n][0];
        a[n][1] = an[n][1];
      }
      if constexpr (w == B_LD) fa_cur = fa_nxt;   // requested a sweep ago, like alpha_{t-1}
    } else if constexpr (w == B_F) {
      scaled = __any_sync(0xffffffffu, need);   // a vote, no branch inside the sweep
      if constexpr (EM) {
        if (on && q == 0 && h == 0) {
          hvec[row0 + t_cur] = fscale;
          B.rn_out[row0 + t_cur] = pinv;
        }
      }
    }
  };
  auto side = [&](auto sc) {
    constexpr int s = decltype(sc)::value;
    run_items_in<0, NIPGPU_PAIR_BWD_A_END, WA, s>(item_a);
    if constexpr (s == MID) mid();
    run_items_in<NIPGPU_PAIR_BWD_B_START, NIPGPU_PAIR_BWD_B_END, WB, s>(item_b);
  };

  {
    const bool p0 = Tw >= 2 && Tw - 2 < T;
    load_half(C.lam_comb + (long long)c_pre * SP, p0, lam);
    if (!EM) load_half(alpha + (row0 + Tw - 2) * SP, p0, an);
    fa_nxt = p0 ? B.fexp[row0 + Tw - 2] : 0;
    fa_pre = ldg_pinned_int(B.fexp + row0 + Tw - 3, Tw >= 3 && Tw - 3 < T);
    if (Tw >= 3 && Tw - 3 < T) c_pre = __ldg(cfg + Tw - 3);
  }
  for (int t = Tw - 1; t >= 1; t--) {
    on = t < T;
    first_next = t - 1 == T - 1;     // slice t-1 is the row's last slice: beta = 1 there
    t_cur = t;
    par = t & 1;
    // u = r . A^T  (k = current state, n = previous state), own n-tiles
    team_sweep<NT, W>(u, r, rth, fp, side);
#pragma unroll
    for (int n = 0; n < NH; n++) {
      beta[n][0] = first_next ? 1.0 : u[n][0];
      beta[n][1] = first_next ? 1.0 : u[n][1];
      r[n][0] = first_next ? lam[n][0] : u[n][0] * lam[n][0];
      r[n][1] = first_next ? lam[n][1] : u[n][1] * lam[n][1];
    }
    if (scaled) {   // warp-uniform and rare: an exact power of two per sequence (1 where not needed)
#pragma unroll
      for (int n = 0; n < NH; n++) {
        beta[n][0] *= fscale; beta[n][1] *= fscale;
        r[n][0] *= fscale; r[n][1] *= fscale;
      }
    }
    eb = first_next ? 0 : eb + fk;   // exponent of beta_{t-1}
    {  // requests for iteration t-1: lambda_{t-2}, alpha_{t-2}, evidence index of slice t-3
      const bool p2 = t >= 2 && t - 2 < T;
      load_half(C.lam_comb + (long long)c_pre * SP, p2, lam);
      if (!EM) load_half(alpha + (row0 + t - 2) * SP, p2, an);
      fa_nxt = fa_pre;   // requested a whole iteration ago (a 4-byte load behind the alpha rows' HBM reads takes that long)
      fa_pre = ldg_pinned_int(B.fexp + row0 + t - 3, t >= 3 && t - 3 < T);
      if (t >= 3 && t - 3 < T) c_pre = __ldg(cfg + t - 3);
    }
  }
  if (Tw >= 1) {   // slice 0: its posterior, and for the E-step r_0 / (phi0 . r_0)
    on = 0 < T;
    t_cur = 0;
    par = 0;
    static_for<0, WA>(item_a);
    double zpart = 0;
    if constexpr (EM) {   // one more record slot: phi0 . r_0 (own half)
      double z0 = 0, z1 = 0;
#pragma unroll
      for (int n = 0; n < NH; n++) {
        z0 += C.phi0[8 * (t0 + n) + 2 * q] * r[n][0];
        z1 += C.phi0[8 * (t0 + n) + 2 * q + 1] * r[n][1];
      }
      zpart = z0 + z1;
      xw(par)[NH << 5] = make_double2(zpart, 0.0);
    }
    mid();
    static_for<B_ST, B_LD>(item_b);
    if constexpr (EM) {
      if (on && q == 0 && h == 0) B.rn_out[row0] = pinv;
      double z = 0;   // phi0 . r_0 from the parts' partials, in the order of the parts
#pragma unroll
      for (int part = 0; part < W; part++) z = part == 0 ? xr(par, part)[NH << 5].x : z + xr(par, part)[NH << 5].x;
      const double zinv = safe_rcp(quad_sum_full(z));
      if (0 < T) {
        double2* out0 = reinterpret_cast<double2*>(r0 + (long long)orig * SP);
#pragma unroll
        for (int n = 0; n < NH; n++) out0[4 * (t0 + n) + q] = make_double2(r[n][0] * zinv, r[n][1] * zinv);
      }
    }
  }
}
