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
#include "jtree.cuh"

#include <cfloat>

namespace nipgpu {

namespace {

struct Work {
  double *tab, *msg, *tmp, *va, *scr;
};

__device__ __forceinline__ Work carve(const DProgram& P, double* W) {
  Work w;
  w.tab = W;
  w.msg = w.tab + P.tab_total;
  w.tmp = w.msg + P.msg_total;
  w.va = w.tmp + P.msg_max;
  w.scr = w.va + 3 * P.S;
  return w;
}

__device__ __forceinline__ void load_tables(const DProgram& P, double* tab, const double* src) {
  for (int i = threadIdx.x; i < P.tab_total; i += blockDim.x) tab[i] = src[i];
  __syncthreads();
}

// dst[j] = sum_r T[base[j] + off[r]] — `lanes` threads share one destination
// entry and combine with a fixed shuffle tree, so the result is deterministic.
__device__ void op_marg(const DProgram& P, const double* tab, int pj, double* dst) {
  const DProj p = P.projs[pj];
  const double* T = tab + p.tab;
  const int* base = P.ipool + p.base;
  const int* off = P.ipool + p.off;
  const int lanes = p.lanes, per_round = blockDim.x / lanes;
  const int sub = threadIdx.x % lanes, jj = threadIdx.x / lanes;
  for (int j0 = 0; j0 < p.m; j0 += per_round) {
    const int j = j0 + jj;
    double s = 0;
    if (j < p.m) {
      const int b = base[j];
      for (int r = sub; r < p.R; r += lanes) s += T[b + off[r]];
    }
    for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sub == 0 && j < p.m) dst[j] = s;
  }
  __syncthreads();
}

// T[base[j] + off[r]] *= v[j]
__device__ void op_absorb(const DProgram& P, double* tab, int pj, const double* v) {
  const DProj p = P.projs[pj];
  double* T = tab + p.tab;
  const int* base = P.ipool + p.base;
  const int* off = P.ipool + p.off;
  const int n = p.m * p.R;
  if (p.lanes == 1) {  // destination holds the fastest dimension: j fastest
    for (int x = threadIdx.x; x < n; x += blockDim.x) {
      const int r = x / p.m, j = x - r * p.m;
      T[base[j] + off[r]] *= v[j];
    }
  } else {
    for (int x = threadIdx.x; x < n; x += blockDim.x) {
      const int j = x / p.R, r = x - j * p.R;
      T[base[j] + off[r]] *= v[j];
    }
  }
  __syncthreads();
}

// T[..] = T[..] * num[j] / den[j], and 0 where den[j] == 0
__device__ void op_absorb_ratio(const DProgram& P, double* tab, int pj, const double* num,
                                const double* den) {
  const DProj p = P.projs[pj];
  double* T = tab + p.tab;
  const int* base = P.ipool + p.base;
  const int* off = P.ipool + p.off;
  const int n = p.m * p.R;
  const bool jfast = p.lanes == 1;
  for (int x = threadIdx.x; x < n; x += blockDim.x) {
    int j, r;
    if (jfast) { r = x / p.m; j = x - r * p.m; } else { j = x / p.R; r = x - j * p.R; }
    const double d = den[j];
    double* e = T + base[j] + off[r];
    *e = (d != 0) ? (*e * num[j]) / d : 0.0;
  }
  __syncthreads();
}

// hard observation: keep only the entries whose state of the variable is `state`
__device__ void op_evidence(const DProgram& P, double* tab, int pj, int state) {
  const DProj p = P.projs[pj];
  double* T = tab + p.tab;
  const int* base = P.ipool + p.base;
  const int* off = P.ipool + p.off;
  const int n = p.m * p.R;
  const bool jfast = p.lanes == 1;
  for (int x = threadIdx.x; x < n; x += blockDim.x) {
    int j, r;
    if (jfast) { r = x / p.m; j = x - r * p.m; } else { j = x / p.R; r = x - j * p.R; }
    if (j != state) T[base[j] + off[r]] = 0.0;
  }
  __syncthreads();
}

__device__ double vec_sum(const double* v, int n, double* red) {
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  return block_sum(s, red);
}

__device__ void vec_normalise(double* v, int n, double* red) {
  const double s = vec_sum(v, n, red);
  if (s != 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] /= s;
  __syncthreads();
}

// collect: child -> parent messages in post-order; sepsets start at 1 so the
// absorbed ratio is the message itself.  Messages are kept for distribute.
__device__ void do_collect(const DProgram& P, const Work& w) {
  for (int i = 0; i < P.n_collect; i++) {
    const DMsg m = P.collect[i];
    op_marg(P, w.tab, m.proj_src, w.msg + m.slot);
    op_absorb(P, w.tab, m.proj_dst, w.msg + m.slot);
  }
}

// distribute: parent -> child; the child absorbs new/old where old is the
// message it sent up during collect.
__device__ void do_distribute(const DProgram& P, const Work& w, const DMsg* list, int n) {
  for (int i = 0; i < n; i++) {
    const DMsg m = list[i];
    op_marg(P, w.tab, m.proj_src, w.tmp);
    op_absorb_ratio(P, w.tab, m.proj_dst, w.tmp, w.msg + m.slot);
    for (int k = threadIdx.x; k < m.size; k += blockDim.x) w.msg[m.slot + k] = w.tmp[k];
    __syncthreads();
  }
}

// returns how many observations were entered
__device__ int enter_row(const DProgram& P, const Work& w, const int* obs, int n_obs,
                         const int* obs_proj) {
  int n = 0;
  for (int k = 0; k < n_obs; k++) {
    const int pj = obs_proj[k], o = obs[k];
    if (pj >= 0 && o >= 0) { op_evidence(P, w.tab, pj, o); n++; }
  }
  return n;
}

__device__ void write_queries(const DProgram& P, const Work& w, const DQuery& Q, double* row,
                              double* red) {
  for (int q = 0; q < Q.n_query; q++) {
    const int pj = Q.proj[q], m = P.projs[pj].m;
    op_marg(P, w.tab, pj, w.scr);
    vec_normalise(w.scr, m, red);
    for (int i = threadIdx.x; i < m; i += blockDim.x) row[Q.off[q] + i] = w.scr[i];
    __syncthreads();
  }
}

__global__ void k_jt_forward(DProgram P, DBatch B, DQuery Q, double* gwork, size_t wstride,
                             int want_ll, int emit, double* alpha, double* post, double* ll_out,
                             int* status_out) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  const Work w = carve(P, gwork ? gwork + (size_t)blockIdx.x * wstride : smem);
  double* vprev = w.va;
  double* vcur = w.va + P.S;
  for (int seq = blockIdx.x; seq < B.n_series; seq += gridDim.x) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    double ll = 0;
    int bad = 0;
    for (int t = 0; t < T; t++) {
      load_tables(P, w.tab, t == 0 ? P.base0 : P.base1);
      double m1 = 0;
      if (t > 0 && P.nif > 0) op_absorb(P, w.tab, P.proj_in, vprev);
      if (want_ll) {
        if (t == 0) m1 = *P.m1_0;
        else if (P.nif == 0) m1 = P.R1[0];
        else {
          double s = 0;
          for (int i = threadIdx.x; i < P.S; i += blockDim.x) s += vprev[i] * P.R1[i];
          m1 = block_sum(s, red);
        }
      }
      const int entered = enter_row(P, w, B.obs + (row0 + t) * B.n_obs, B.n_obs, B.obs_proj);
      do_collect(P, w);
      double m2 = 0;
      if (want_ll) m2 = vec_sum(w.tab + P.root_tab, P.root_size, red);
      // a slice without any evidence has m2 == m1 by definition; do not let rounding decide
      // whether the running log-likelihood is "> 0" (the reference's BAD_LUCK test)
      if (want_ll && entered == 0) m2 = m1;
      if (emit) {
        do_distribute(P, w, P.distribute, P.n_distribute);
        if (post) write_queries(P, w, Q, post + (row0 + t) * Q.row, red);
      } else if (P.nif > 0)
        do_distribute(P, w, P.path, P.n_path);
      if (P.nif > 0) {
        op_marg(P, w.tab, P.proj_out, vcur);
        vec_normalise(vcur, P.S, red);
        if (alpha)
          for (int i = threadIdx.x; i < P.S; i += blockDim.x) alpha[(row0 + t) * P.S + i] = vcur[i];
        double* x = vprev; vprev = vcur; vcur = x;
        __syncthreads();
      }
      if (want_ll) {  // src/nip.c:1458-1474 and the BAD_LUCK test of e_step, :1827-1831
        if (m1 > 0 && m2 > 0) ll += log(m2) - log(m1);
        if (m2 == 0) ll = -DBL_MAX;
        if (m1 <= 0 || m2 <= 0 || ll > 0) bad = 1;
      }
    }
    if (threadIdx.x == 0) {
      if (ll_out) ll_out[seq] = ll;
      if (status_out) status_out[seq] = bad;
    }
  }
}

__global__ void k_jt_backward(DProgram P, DBatch B, DQuery Q, double* gwork, size_t wstride,
                              const double* alpha, double* post, double* acc, long long acc_stride) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  const Work w = carve(P, gwork ? gwork + (size_t)blockIdx.x * wstride : smem);
  double* a_prev = w.va;          // alpha_{t-1}
  double* a_cur = w.va + P.S;     // alpha_t
  double* gam = w.va + 2 * P.S;   // gamma_{t+1}
  double* my_acc = acc ? acc + (size_t)blockIdx.x * acc_stride : nullptr;
  for (int seq = blockIdx.x; seq < B.n_series; seq += gridDim.x) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    for (int t = T - 1; t >= 0; t--) {
      load_tables(P, w.tab, t == 0 ? P.base0 : P.base1);
      if (t > 0 && P.nif > 0) {
        for (int i = threadIdx.x; i < P.S; i += blockDim.x) a_prev[i] = alpha[(row0 + t - 1) * P.S + i];
        __syncthreads();
        op_absorb(P, w.tab, P.proj_in, a_prev);
      }
      enter_row(P, w, B.obs + (row0 + t) * B.n_obs, B.n_obs, B.obs_proj);
      if (t < T - 1 && P.nif > 0) op_absorb_ratio(P, w.tab, P.proj_out, gam, a_cur);
      do_collect(P, w);
      do_distribute(P, w, P.distribute, P.n_distribute);
      if (post) write_queries(P, w, Q, post + (row0 + t) * Q.row, red);
      if (my_acc) {  // e_step "THE CORE", src/nip.c:1925-1967
        for (int v = 0; v < P.nv; v++) {
          if (t > 0 && (P.var_flags[v] & NIPGPU_IF_OLD_OUTGOING)) continue;
          const int pj = P.proj_fam[v], m = P.projs[pj].m;
          op_marg(P, w.tab, pj, w.scr);
          const double tot = vec_sum(w.scr, m, red);
          if (tot != 0)
            for (int i = threadIdx.x; i < m; i += blockDim.x) my_acc[P.coff[v] + i] += w.scr[i] / tot;
          __syncthreads();
        }
      }
      if (t > 0 && P.nif > 0) {
        op_marg(P, w.tab, P.proj_in, gam);
        vec_normalise(gam, P.S, red);
      }
      double* x = a_prev; a_prev = a_cur; a_cur = x;
      __syncthreads();
    }
  }
}

__global__ void k_jt_likelihood(DProgram P, DBatch B, const int* proj_off, const int* proj_on,
                                double* gwork, size_t wstride, double* out) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  const Work w = carve(P, gwork ? gwork + (size_t)blockIdx.x * wstride : smem);
  for (int seq = blockIdx.x; seq < B.n_series; seq += gridDim.x) {
    const int T = B.len[seq];
    const long long row0 = B.row_off[seq];
    for (int t = 0; t < T; t++) {
      const int* obs = B.obs + (row0 + t) * B.n_obs;
      load_tables(P, w.tab, t == 0 ? P.base0 : P.base1);
      enter_row(P, w, obs, B.n_obs, proj_off);
      do_collect(P, w);
      const double m1 = vec_sum(w.tab + P.root_tab, P.root_size, red);
      load_tables(P, w.tab, t == 0 ? P.base0 : P.base1);
      enter_row(P, w, obs, B.n_obs, proj_off);
      const int extra = enter_row(P, w, obs, B.n_obs, proj_on);
      do_collect(P, w);
      double m2 = vec_sum(w.tab + P.root_tab, P.root_size, red);
      if (extra == 0) m2 = m1;
      if (threadIdx.x == 0) { out[(row0 + t) * 2] = m1; out[(row0 + t) * 2 + 1] = m2; }
    }
  }
}

// R1[i] = sum over everything but I_{t-1} of base1, m1_0 = total mass of base0
__global__ void k_jt_calibrate(DProgram P, double* gwork, double* R1, double* m1_0) {
  extern __shared__ double smem[];
  __shared__ double red[40];
  const Work w = carve(P, gwork ? gwork : smem);
  load_tables(P, w.tab, P.base1);
  do_collect(P, w);
  if (P.nif > 0) {
    do_distribute(P, w, P.distribute, P.n_distribute);
    op_marg(P, w.tab, P.proj_in, w.va);
    for (int i = threadIdx.x; i < P.S; i += blockDim.x) R1[i] = w.va[i];
    __syncthreads();
  } else {
    const double s = vec_sum(w.tab + P.root_tab, P.root_size, red);
    if (threadIdx.x == 0) R1[0] = s;
  }
  load_tables(P, w.tab, P.base0);
  do_collect(P, w);
  const double s0 = vec_sum(w.tab + P.root_tab, P.root_size, red);
  if (threadIdx.x == 0) *m1_0 = s0;
}

__global__ void k_jt_slice(DProgram P, double* gwork, const double* start, double* out_tables,
                           double* out_msgs) {
  extern __shared__ double smem[];
  const Work w = carve(P, gwork ? gwork : smem);
  load_tables(P, w.tab, start);
  do_collect(P, w);
  do_distribute(P, w, P.distribute, P.n_distribute);
  for (int i = threadIdx.x; i < P.tab_total; i += blockDim.x) out_tables[i] = w.tab[i];
  for (int i = threadIdx.x; i < P.msg_total; i += blockDim.x) out_msgs[i] = w.msg[i];
}

// single-slice API: mass = sum of cliques - sum of sepsets (nip_probability_mass,
// src/nipjointree.c:1156-1188) of the consistent tables left by k_jt_slice
__global__ void k_jt_mass(const double* tables, int n_tab, const double* msgs, int n_msg, double* out) {
  __shared__ double red[40];
  const double a = vec_sum(tables, n_tab, red);
  const double b = vec_sum(msgs, n_msg, red);
  if (threadIdx.x == 0) *out = a - b;
}

// get_probability (src/nip.c:2261-2298): normalised one-variable marginal of the family clique
__global__ void k_jt_marginal(DProgram P, const double* tables, int pj, double* out) {
  __shared__ double red[40];
  const int m = P.projs[pj].m;
  op_marg(P, tables, pj, out);
  vec_normalise(out, m, red);
}

template <class K>
int prep_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    NIPGPU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NIPGPU_OK;
}

}  // namespace

int jt_forward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l, int want_ll,
               int emit, double* alpha, double* post, double* ll, int* status, cudaStream_t st) {
  if (int e = prep_smem(k_jt_forward, l.smem_bytes)) return e;
  k_jt_forward<<<l.grid, l.threads, l.smem_bytes, st>>>(p, b, q, l.gwork, jt_work_doubles(p), want_ll,
                                                        emit, alpha, post, ll, status);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_backward(const DProgram& p, const DBatch& b, const DQuery& q, const JtLaunch& l,
                const double* alpha, double* post, double* acc, long long acc_stride, cudaStream_t st) {
  if (int e = prep_smem(k_jt_backward, l.smem_bytes)) return e;
  k_jt_backward<<<l.grid, l.threads, l.smem_bytes, st>>>(p, b, q, l.gwork, jt_work_doubles(p), alpha,
                                                         post, acc, acc_stride);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_likelihood(const DProgram& p, const DBatch& b, const int* proj_off, const int* proj_on,
                  const JtLaunch& l, double* out, cudaStream_t st) {
  if (int e = prep_smem(k_jt_likelihood, l.smem_bytes)) return e;
  k_jt_likelihood<<<l.grid, l.threads, l.smem_bytes, st>>>(p, b, proj_off, proj_on, l.gwork,
                                                           jt_work_doubles(p), out);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

int jt_calibrate(const DProgram& p, const JtLaunch& l, double* R1, double* m1_0, cudaStream_t st) {
  if (int e = prep_smem(k_jt_calibrate, l.smem_bytes)) return e;
  k_jt_calibrate<<<1, l.threads, l.smem_bytes, st>>>(p, l.gwork, R1, m1_0);
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
  if (int e = prep_smem(k_jt_slice, l.smem_bytes)) return e;
  k_jt_slice<<<1, l.threads, l.smem_bytes, st>>>(p, l.gwork, start, out_tables, out_msgs);
  NIPGPU_LAUNCHED();
  return NIPGPU_OK;
}

}  // namespace nipgpu
