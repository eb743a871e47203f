// sweep.cuh — the [8 sequences x SP] . [SP x SP] contraction of one slice on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; on sm_100a every wider f64 mma shape,
// m16n8k4/k8/k16, lowers to the same instruction, so there is nothing bigger to issue).
// Shared by the chain kernels (chain.cu) and the sweep probe (probe.cu).
#pragma once

#include <type_traits>

namespace nipgpu {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma_init(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};"
               : "=d"(c0), "=d"(c1)
               : "d"(a), "d"(b), "d"(0.0));
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// How the B fragments reach the tensor instructions (measured with nipgpu_probe_sweep, see
// tools/sweep_probe.py and profiles/):
//   0  plain shared loads, fetched one k-step ahead in the source.  ptxas re-orders the fully
//      unrolled sweep into pairs of n-tiles with the k loop inside (two accumulator chains), and
//      re-uses the B registers of a DMMA for the very next LDS (a read-barrier wait per load)
//   1  ld.volatile.shared: the loads keep their program order, so the sweep stays k-step major
//      (eight accumulator chains)
//   2  as 1, fetched two k-steps ahead (three buffers)
//   3  (probe only, wrong arithmetic) no loads after the first k-step: the tensor pipe's own pace
#ifndef NIPGPU_SWEEP_VARIANT
#define NIPGPU_SWEEP_VARIANT 0
#endif

template <int V>
__device__ __forceinline__ double2 lds_b(const double2* p) {
  if constexpr (V == 0) {
    return *p;
  } else {
    double2 v;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
  }
}

// acc[n] = sum over the 2*NT k-steps of A(k-step) x B(k-step, n): 2*NT*NT tensor instructions.
// One 16-byte shared load brings the B fragments of two neighbouring n-tiles.
//
// A warp can issue one DMMA per 16 cycles and is alone on its scheduler, so the issue slots
// between two DMMAs are free.  `side(slot)` is called after every tensor instruction with a
// compile-time slot number 0 .. 2*NT*NT-1; callers hang small pieces of work there that do not
// depend on this sweep's result.
template <int NT, int V = NIPGPU_SWEEP_VARIANT, class Side>
__device__ __forceinline__ void mma_sweep(double (&acc)[NT][2], const double (&a)[NT][2],
                                          const double* __restrict__ frag, Side side) {
  constexpr int AHEAD = V == 2 ? 2 : 1, NB = AHEAD + 1;
  double b[NB][NT];
  auto fetch = [&](auto ksc) {
    constexpr int ks = decltype(ksc)::value, buf = ks % NB;
    if constexpr (NT >= 2) {
      const double2* p = reinterpret_cast<const double2*>(frag) + ((ks * (NT / 2)) << 5);
      static_for<0, NT / 2>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        const double2 v = lds_b<V>(p + (n2 << 5));
        b[buf][2 * n2] = v.x;
        b[buf][2 * n2 + 1] = v.y;
      });
    } else {
      b[buf][0] = frag[ks << 5];
    }
  };
  static_for<0, (V == 3 ? NB : AHEAD)>([&](auto kc) { fetch(kc); });
  static_for<0, 2 * NT>([&](auto ksc) {
    constexpr int ks = decltype(ksc)::value;
    if constexpr (ks + AHEAD < 2 * NT && V != 3) fetch(std::integral_constant<int, ks + AHEAD>{});
    const double av = a[ks >> 1][ks & 1];
    static_for<0, NT>([&](auto nc) {
      constexpr int n = decltype(nc)::value;
      if constexpr (ks == 0) dmma_init(acc[n][0], acc[n][1], av, b[0][n]);
      else dmma(acc[n][0], acc[n][1], av, b[ks % NB][n]);
      side(std::integral_constant<int, ks * NT + n>{});
    });
  });
}

// Spreads W work items evenly over the NS side slots of a sweep: slot s runs the items
// w with  w*NS/W == s  (several per slot when W > NS).  Everything folds at compile time.
template <int NS, int W, int SLOT, class Item>
__device__ __forceinline__ void run_items(Item& item) {
  static_for<0, W>([&](auto wc) {
    if constexpr ((decltype(wc)::value * NS) / W == SLOT) item(wc);
  });
}

}  // namespace nipgpu
