// api.cu — implementation of the C ABI declared in include/nipgpu.h.
#include "api.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace nipgpu {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

namespace {

template <class T>
int dev_upload(T** dst, const T* src, size_t n, cudaStream_t st) {
  NIPGPU_CUDA(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) NIPGPU_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return NIPGPU_OK;
}

template <class T>
int dev_upload(T** dst, const std::vector<T>& v, cudaStream_t st) {
  return dev_upload(dst, v.data(), v.size(), st);
}

int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

// ---- launch geometry of the generic engine ------------------------------
int choose_launch(nipgpu_model* m) {
  const HostModel& hm = m->hm;
  int biggest = 1;
  for (int c = 0; c < hm.nc; c++) biggest = std::max(biggest, hm.csize[c]);
  JtLaunch& l = m->launch;
  l = JtLaunch{};
  size_t work = jt_work_doubles(m->prog);
  const size_t bytes = work * sizeof(double);
  // NIPGPU_JT_MODE = warp | cta | hbm | grid overrides the choice below; the tests use it to run
  // every mode on small models (NIPGPU_FORCE_HBM_WORKSPACE=1 is the older spelling of "hbm")
  const char* env = getenv("NIPGPU_JT_MODE");
  std::string want = env ? env : "";
  const char* force = getenv("NIPGPU_FORCE_HBM_WORKSPACE");
  if (want.empty() && force && force[0] == '1') want = "hbm";
  if (want.empty()) {
    if (biggest <= 64 && bytes <= 8 * 1024) want = "warp";
    else if (bytes <= 200 * 1024) want = "cta";
    // measured (600 sequences, tools/dev_midsize.py): 7^6-entry cliques 2.5e4 slice-steps/s on per-CTA
    // HBM workspaces vs 1.9e4 with the grid team's concurrent groups; 8^6: 1.2e4 vs 1.6e4
    else if (biggest >= (1 << 18)) want = "grid";
    else want = "hbm";
  }
  if ((want == "warp" && bytes > 24 * 1024) || (want == "cta" && bytes > 200 * 1024))
    return fail(NIPGPU_EINVAL, "NIPGPU_JT_MODE: the tables of this model do not fit shared memory");
  auto need_gwork = [&](size_t doubles) {
    if (m->gwork_doubles >= doubles) return NIPGPU_OK;
    cudaFree(m->d_gwork);
    m->d_gwork = nullptr;
    m->gwork_doubles = 0;
    NIPGPU_CUDA(cudaMalloc((void**)&m->d_gwork, doubles * sizeof(double)));
    m->gwork_doubles = doubles;
    return NIPGPU_OK;
  };
  if (want == "warp") {
    l.mode = JT_MODE_WARP;
    int wpc = 8;
    while (wpc > 1 && wpc * bytes > 64 * 1024) wpc /= 2;
    l.threads = 32 * wpc;
    l.smem_bytes = wpc * bytes;
    const int by_smem = (int)std::max<size_t>(1, (200 * 1024) / (l.smem_bytes + 1024));
    l.grid = m->sm_count * std::max(1, std::min({by_smem, 2048 / l.threads, 32}));
  } else if (want == "cta") {
    l.mode = JT_MODE_CTA;
    l.threads = biggest <= 64 ? 32 : biggest <= 256 ? 64 : biggest <= 1024 ? 128 : 256;
    l.smem_bytes = bytes;
    const int by_smem = (int)std::max<size_t>(1, (220 * 1024) / (bytes + 1024));
    // a single CTA per SM (its tables fill the shared memory): make it a big one, the index-map
    // gathers of the potential operations need warps to hide behind
    if (by_smem == 1 && biggest >= 2048) l.threads = 512;
    l.grid = m->sm_count * std::max(1, std::min({by_smem, 2048 / l.threads, 16}));
  } else if (want == "grid") {
    l.mode = JT_MODE_GRID;
    l.threads = 256;
    work = jt_work_doubles(m->prog, true);
    const int all_ctas = jt_grid_ctas(l.threads, m->sm_count, &l.smem_bytes);
    // Several sequences side by side, each streamed by its own cooperative kernel on a share of
    // the SMs: with tables of a few MB a slice is barrier-latency bound (four 16^5-entry cliques:
    // 2.0e3 -> 4.6e3 slice-steps/s), and even the 403 MB of C3 gain (433 -> 589) because one
    // group's barriers and short operations hide under the others' streaming passes
    int groups = 1;
    while (groups < 8 && 2 * groups * bytes <= ((size_t)16 << 30) && all_ctas / (2 * groups) >= 32) groups *= 2;
    const char* eg = getenv("NIPGPU_JT_GROUPS");
    if (eg && atoi(eg) >= 1 && atoi(eg) <= 8) groups = atoi(eg);
    l.groups = groups;
    l.grid = all_ctas / groups;
    // Concurrent cooperative kernels are only validated one by one by the runtime: together they
    // must never ask for more CTAs than can be resident, or each would spin in its grid barrier
    // waiting for CTAs the others keep out.  All groups run the same kernel (same occupancy) and
    // join before the next kind is launched; a spare slot per group leaves room for a stray CTA.
    if (groups > 1 && l.grid > 8) l.grid -= 1;
    const size_t part = 2 * (size_t)l.grid + 8, scratch = (size_t)l.grid * l.threads;
    l.group_stride = (work + part + scratch + 1) & ~(size_t)1;   // keeps 16-byte alignment
    if (int e = need_gwork(l.group_stride * groups)) return e;
    l.gwork = m->d_gwork;
    l.part = m->d_gwork + work;
    l.scratch = l.part + part;
    for (int g = 0; g < groups; g++) {
      if (g + 1 < groups && !m->aux_stream[g]) NIPGPU_CUDA(cudaStreamCreateWithFlags(&m->aux_stream[g], cudaStreamNonBlocking));
      if (!m->aux_event[g]) NIPGPU_CUDA(cudaEventCreateWithFlags(&m->aux_event[g], cudaEventDisableTiming));
    }
    for (int g = 0; g < 8; g++) l.aux_stream[g] = m->aux_stream[g];
    for (int g = 0; g < 9; g++) l.aux_event[g] = m->aux_event[g];
    const char* tr = getenv("NIPGPU_JT_TRACE");
    if (tr && tr[0] == '1' && !m->d_trace) {
      NIPGPU_CUDA(cudaMalloc((void**)&m->d_trace, JT_TRACE_WORDS * sizeof(unsigned long long)));
      NIPGPU_CUDA(cudaMemset(m->d_trace, 0, JT_TRACE_WORDS * sizeof(unsigned long long)));
    }
    l.trace = m->d_trace;
  } else {
    l.mode = JT_MODE_CTA;
    l.threads = biggest <= 64 ? 32 : biggest <= 256 ? 64 : biggest <= 1024 ? 128 : 256;
    size_t ctas = (size_t)m->sm_count * 2;
    const size_t budget = (size_t)32 << 30;  // keep the workspace under 32 GB
    while (ctas > 1 && ctas * bytes > budget) ctas /= 2;
    if (int e = need_gwork(ctas * work)) return e;
    l.gwork = m->d_gwork;
    l.grid = (int)ctas;
  }
  l.slots = l.mode == JT_MODE_WARP ? l.grid * (l.threads / 32) : l.mode == JT_MODE_GRID ? l.groups : l.grid;
  return NIPGPU_OK;
}

// base0/base1, calibration vectors and the chain-engine tables from d_orig / d_prior
int refresh_derived(nipgpu_model* m, bool from_counts = false) {
  const HostModel& hm = m->hm;
  cudaStream_t st = m->stream;
  const size_t tab_bytes = (size_t)m->prog.tab_total * sizeof(double);
  NIPGPU_CUDA(cudaMemcpyAsync(m->d_base1, m->d_orig, tab_bytes, cudaMemcpyDeviceToDevice, st));
  const int np = (int)hm.prior_vars.size();
  if (int e = prior_flags(m->d_prior, m->d_prior_off, m->d_prior_vars, np, m->d_prior_flags, st)) return e;
  auto apply = [&](double* base, int k) {
    const int v = hm.prior_vars[k], c = hm.family[v];
    int stride = 1;
    for (int j = 0; j < hm.var_pos(c, v); j++) stride *= hm.card[hm.clique_vars(c)[j]];
    return apply_vector(base + m->tab_off[c], hm.csize[c], stride, hm.card[v],
                        m->d_prior + hm.prior_off[v], m->d_prior_flags + k, st);
  };
  for (int k = 0; k < np; k++)  // priors kept when the slice has history (src/nip.c:100)
    if (!(hm.flags[hm.prior_vars[k]] & NIPGPU_IF_OLD_OUTGOING))
      if (int e = apply(m->d_base1, k)) return e;
  NIPGPU_CUDA(cudaMemcpyAsync(m->d_base0, m->d_base1, tab_bytes, cudaMemcpyDeviceToDevice, st));
  for (int k = 0; k < np; k++)
    if (hm.flags[hm.prior_vars[k]] & NIPGPU_IF_OLD_OUTGOING)
      if (int e = apply(m->d_base0, k)) return e;
  if (m->launch.threads == 0) return fail(NIPGPU_EINVAL, "launch not configured");
  if (int e = jt_calibrate(m->prog, m->launch, m->d_R1, m->d_m10, st)) return e;
  if (m->chain.ok)
    if (int e = chain_refresh(hm, m->chain, m->d_base0, m->d_base1, m->tab_off, m->d_ipool, m->d_R1, m->d_m10, st)) return e;
  if (m->fac.ok)   // engine 3: its factors are the CPTs themselves (after an M-step: the normalised counts)
    if (int e = fac_refresh(hm, m->fac, m->d_prior, m->d_prior_flags, from_counts ? m->d_counts : nullptr, st)) return e;
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  m->slice_consistent = false;
  return NIPGPU_OK;
}

int ensure_post(nipgpu_batch* b, size_t doubles) {
  if (b->post_cap >= doubles) return NIPGPU_OK;
  cudaFree(b->d_post);
  b->d_post = nullptr;
  b->post_cap = 0;
  NIPGPU_CUDA(cudaMalloc((void**)&b->d_post, std::max<size_t>(doubles, 1) * sizeof(double)));
  b->post_cap = doubles;
  return NIPGPU_OK;
}

DBatch dev_batch(const nipgpu_batch* b, const int* obs_proj) {
  DBatch d;
  d.n_series = b->n_series; d.n_obs = b->n_obs; d.len = b->d_len; d.row_off = b->d_row_off;
  d.obs = b->d_obs; d.obs_proj = obs_proj;
  return d;
}

// evidence columns -> projection ids (or -1) in slot `which` of the batch scratch
int upload_obs_proj(nipgpu_model* m, nipgpu_batch* b, const uint8_t* mask, bool null_means_all,
                    int which, const int** out) {
  std::vector<int> p(std::max(b->n_obs, 1), -1);
  for (int k = 0; k < b->n_obs; k++) {
    const int v = b->obs_vars[k];
    const bool on = mask ? mask[v] != 0 : null_means_all;
    p[k] = on ? m->hm.proj_var[v] : -1;
  }
  int* dst = b->d_obs_proj + (size_t)which * std::max(b->n_obs, 1);
  NIPGPU_CUDA(cudaMemcpyAsync(dst, p.data(), p.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));  // `p` dies here
  *out = dst;
  return NIPGPU_OK;
}

int upload_query(nipgpu_model* m, nipgpu_batch* b, int nq, const int32_t* q, DQuery* out) {
  const HostModel& hm = m->hm;
  std::vector<int> proj(std::max(nq, 1), 0), off(std::max(nq, 1), 0);
  int row = 0;
  for (int i = 0; i < nq; i++) {
    if (q[i] < 0 || q[i] >= hm.nv) return fail(NIPGPU_EINVAL, "query variable out of range");
    proj[i] = hm.proj_var[q[i]];
    off[i] = row;
    row += hm.card[q[i]];
  }
  if (b->q_cap < (size_t)std::max(nq, 1)) {
    cudaFree(b->d_qproj); cudaFree(b->d_qoff);
    b->q_cap = std::max(nq, 1);
    NIPGPU_CUDA(cudaMalloc((void**)&b->d_qproj, b->q_cap * sizeof(int)));
    NIPGPU_CUDA(cudaMalloc((void**)&b->d_qoff, b->q_cap * sizeof(int)));
  }
  NIPGPU_CUDA(cudaMemcpyAsync(b->d_qproj, proj.data(), proj.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NIPGPU_CUDA(cudaMemcpyAsync(b->d_qoff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  out->n_query = nq; out->row = row; out->proj = b->d_qproj; out->off = b->d_qoff;
  return NIPGPU_OK;
}

int ensure_alpha(nipgpu_model* m, nipgpu_batch* b) {
  if (b->d_alpha) return NIPGPU_OK;
  NIPGPU_CUDA(cudaMalloc((void**)&b->d_alpha,
                         std::max<size_t>((size_t)b->rows * m->hm.S, 1) * sizeof(double)));
  return NIPGPU_OK;
}

void fill_fac_args(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, FacRunArgs* a) {
  a->n_series = b->n_series; a->n_obs = b->n_obs; a->t_max = b->t_max; a->rows = b->rows;
  a->len = &b->len; a->row_off = &b->row_off; a->obs_vars = &b->obs_vars; a->use_evidence = use_evidence;
  a->d_obs = b->d_obs; a->d_row_off = b->d_row_off; a->d_alpha = b->d_alpha; a->d_ll = b->d_ll;
  a->d_status = b->d_status; a->d_R1 = m->d_R1; a->d_m10 = m->d_m10;
}

// NIPGPU_CHUNKED_COPY_MIN_MB: smallest posterior set that is copied in chunks (default 256 MB)
static size_t chunked_copy_min_bytes() {
  static const size_t v = [] {
    const char* p = getenv("NIPGPU_CHUNKED_COPY_MIN_MB");
    return (size_t)(p && atoi(p) > 0 ? atoi(p) : 256) << 20;
  }();
  return v;
}

// host_post != nullptr: the caller wants the posterior rows in host memory; when the batch allows it
// (see below) they are copied chunk by chunk while later chunks are still being computed, and
// *copied is set
int infer_impl(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int nq,
               const int32_t* query, int forward_only, int want_ll, bool want_post,
               double* host_post = nullptr, bool* copied = nullptr) {
  if (!m || !b || b->m != m) return fail(NIPGPU_EINVAL, "model/batch mismatch");
  const HostModel& hm = m->hm;
  DQuery Q{};
  if (int e = upload_query(m, b, nq, query, &Q)) return e;
  if (want_post)
    if (int e = ensure_post(b, (size_t)b->rows * std::max(Q.row, 1))) return e;
  double* post = want_post && nq > 0 ? b->d_post : nullptr;
  m->last_kernel_ms = 0;
  m->last_kernel_n = 0;

  // ---- engine 2 when the model, the evidence columns and the query allow it ----
  // The chain engines produce the posterior of the JOINT interface state.  A query for the single
  // interface variable is that vector itself; every other variable of a chain-structured model
  // (variables of a composite interface, previous-slice interface variables, the variables of
  // the leaf cliques) is a linear function of it, evaluated by chain_post_vars afterwards.
  ChainPlan plan;
  std::vector<ChainQueryVar> qv;
  const bool direct = nq == 0 || (nq == 1 && hm.nif == 1 && query[0] == hm.outg[0]);
  const bool plan_ok = m->engine == NIPGPU_ENGINE_CHAIN && m->chain.ok &&
                       chain_plan(hm, m->chain, b->n_obs, b->obs_vars.data(), use_evidence, plan);
  if (plan_ok && (direct || !post || chain_query_plan(hm, m->chain, plan, nq, query, forward_only, qv))) {
    if (int e = chain_batch_prepare(m->chain, b->chain, b->n_series, b->len.data(), b->rows, b->t_max, m->stream)) return e;
    const bool project = !direct && post != nullptr;
    const int SPc = m->chain.SP;
    if (project && !b->d_joint)
      NIPGPU_CUDA(cudaMalloc((void**)&b->d_joint, std::max<size_t>((size_t)b->rows * SPc, 1) * sizeof(double)));
    if (project && !b->d_first) {
      NIPGPU_CUDA(cudaMalloc((void**)&b->d_first, std::max<long long>(b->rows, 1)));
      NIPGPU_CUDA(cudaMemsetAsync(b->d_first, 0, std::max<long long>(b->rows, 1), m->stream));
      if (int e = jt_first_rows(b->d_row_off, b->n_series, b->rows, b->d_first, m->stream)) return e;
    }
    ChainInferArgs a;
    a.n_series = b->n_series; a.n_obs = b->n_obs; a.t_max = b->t_max; a.rows = b->rows;
    a.d_obs = b->d_obs; a.d_row_off = b->d_row_off; a.want_ll = want_ll; a.forward_only = forward_only;
    a.d_post = project ? b->d_joint : post; a.post_stride = project ? SPc : Q.row; a.post_off = 0;
    a.d_ll = b->d_ll; a.d_status = b->d_status;
    NIPGPU_CUDA(cudaEventRecord(m->ev_mid, m->stream));   // re-recorded between the two kernels of the warp-resident path
    // Host-buffered smoothing of a large set is a PCIe transfer with some kernels in front
    // (C2: 2.9 ms of kernels, 40 ms of device-to-host copy).  When the series are stored in
    // length-sorted order (equal lengths: always) a range of sorted positions is a contiguous
    // block of output rows, so the set runs in chunks and chunk k's rows travel on a second
    // stream while chunk k+1 is computed.
    int n_chunks = 1;
    if (host_post && copied && !project && post && !forward_only && m->chain.NT > 1 && !m->chain.dense &&
        b->n_series >= 2048 && (size_t)b->rows * Q.row * sizeof(double) >= chunked_copy_min_bytes()) {
      bool identity = true;
      for (int i = 0; i < b->n_series && identity; i++) identity = b->chain.order[i] == i;
      static const bool off = [] { const char* p = getenv("NIPGPU_CHUNKED_COPY"); return p && p[0] == '0'; }();
      if (identity && !off) n_chunks = 4;
    }
    if (n_chunks > 1) {
      if (!m->copy_stream) {
        NIPGPU_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) NIPGPU_CUDA(cudaEventCreateWithFlags(&m->chunk_ev[k], cudaEventDisableTiming));
      }
      const int per = ((b->n_series + n_chunks - 1) / n_chunks + 31) / 32 * 32;
      cudaEvent_t first_ev = m->ev0;
      for (int k = 0; k < n_chunks; k++) {
        a.p0 = std::min(k * per, b->n_series);
        a.p1 = std::min(a.p0 + per, b->n_series);
        if (a.p1 <= a.p0) break;
        if (int ce = chain_infer(hm, m->chain, b->chain, plan, a, m->stream, first_ev, m->ev1)) return ce;
        first_ev = nullptr;
        NIPGPU_CUDA(cudaEventRecord(m->chunk_ev[k], m->stream));
        NIPGPU_CUDA(cudaStreamWaitEvent(m->copy_stream, m->chunk_ev[k], 0));
        const long long r0 = b->row_off[a.p0];
        const long long r1 = a.p1 < b->n_series ? b->row_off[a.p1] : b->rows;
        NIPGPU_CUDA(cudaMemcpyAsync(host_post + r0 * Q.row, post + r0 * Q.row, (size_t)(r1 - r0) * Q.row * sizeof(double),
                                    cudaMemcpyDeviceToHost, m->copy_stream));
      }
      NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
      NIPGPU_CUDA(cudaStreamSynchronize(m->copy_stream));
      *copied = true;
      float ms = 0;
      if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) { m->last_kernel_ms = ms; m->last_kernel_n = 2 * n_chunks; }
      return NIPGPU_OK;
    }
    g_chain_mid_event = m->ev_mid;
    const int ce = chain_infer(hm, m->chain, b->chain, plan, a, m->stream, m->ev0, m->ev1);
    g_chain_mid_event = nullptr;
    if (ce) return ce;
    if (project)
      if (int e = chain_post_vars(hm, m->chain, b->chain, qv, m->d_base0, m->d_base1, m->tab_off, m->d_ipool,
                                  b->d_joint, b->d_first, b->d_row_off, b->n_series, b->rows, Q.row, post, m->stream))
        return e;
    NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
    float ms = 0;
    if (b->n_series > 0 && cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) {
      m->last_kernel_ms = ms;
      m->last_kernel_n = forward_only ? 1 : 2;
      float f = 0;
      m->last_forward_ms = cudaEventElapsedTime(&f, m->ev0, m->ev_mid) == cudaSuccess ? f : 0;
    }
    return NIPGPU_OK;
  }

  // ---- engine 3: factor by factor ----
  if (m->engine == NIPGPU_ENGINE_FACTOR && m->fac.ok) {
    FacRunArgs a{};
    fill_fac_args(m, b, use_evidence, &a);
    a.n_query = nq; a.query = query; a.post_row = Q.row; a.d_post = post;
    a.want_ll = want_ll; a.forward_only = forward_only;
    NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
    const int fe = fac_run(hm, m->fac, a, m->stream);
    if (fe == NIPGPU_EUNSUPPORTED && m->fac_auto) {
      m->engine = NIPGPU_ENGINE_JTREE;   // a contraction beyond the planner's limits: engine 1 from now on
    } else {
      if (fe) return fe;
      NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
      NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
      float ms = 0;
      if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) { m->last_kernel_ms = ms; m->last_kernel_n = 0; }
      return NIPGPU_OK;
    }
  }

  // ---- generic join-tree engine ----
  const int* obs_proj = nullptr;
  if (int e = upload_obs_proj(m, b, use_evidence, true, 0, &obs_proj)) return e;
  if (int e = ensure_alpha(m, b)) return e;
  const DBatch B = dev_batch(b, obs_proj);
  const JtLaunch l = jt_fit(m->launch, b->n_series);
  NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
  if (int e = jt_forward(m->prog, B, Q, l, want_ll, forward_only, b->d_alpha, forward_only ? post : nullptr,
                         b->d_ll, b->d_status, m->stream)) return e;
  if (!forward_only)
    if (int e = jt_backward(m->prog, B, Q, l, b->d_alpha, post, nullptr, 0, m->stream)) return e;
  NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) {
    m->last_kernel_ms = ms;
    m->last_kernel_n = forward_only ? 1 : 2;
  }
  return NIPGPU_OK;
}

}  // namespace
}  // namespace nipgpu

using namespace nipgpu;

extern "C" {

const char* nipgpu_last_error(void) { return g_error.c_str(); }

int nipgpu_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(NIPGPU_ENODEVICE, "no CUDA device");
  if (device < 0 || device >= n) return fail(NIPGPU_ENODEVICE, "device index out of range");
  cudaDeviceProp p;
  NIPGPU_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return fail(NIPGPU_ENODEVICE, std::string("device is not sm_100-class (Blackwell B200): ") + p.name);
  return NIPGPU_OK;
}

// inside nipgpu_model_create: a CUDA failure releases the half-built model (stream, events, device memory)
#define CREATE_CUDA(call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                       \
      nipgpu_model_destroy(m);                                                              \
      return NIPGPU_ECUDA;                                                                  \
    }                                                                                       \
  } while (0)

int nipgpu_model_create(const nipgpu_model_desc* desc, int device, int engine, nipgpu_model** out) {
  if (!out) return fail(NIPGPU_EINVAL, "null out pointer");
  *out = nullptr;
  if (int e = nipgpu_device_check(device)) return e;
  NIPGPU_CUDA(cudaSetDevice(device));
  nipgpu_model* m = new nipgpu_model();
  m->device = device;
  const std::string err = m->hm.load(desc);
  if (!err.empty()) { delete m; return fail(NIPGPU_EINVAL, "model description: " + err); }
  HostModel& hm = m->hm;
  for (int c = 0; c < hm.nc; c++)
    if (hm.clique_dim(c) > kMaxDims) { delete m; return fail(NIPGPU_EUNSUPPORTED, "clique with too many variables"); }
  const std::string why = chain_build(hm, m->chain);
  if (engine == NIPGPU_ENGINE_CHAIN && !m->chain.ok) { delete m; return fail(NIPGPU_EUNSUPPORTED, "chain engine: " + why); }
  if (engine < NIPGPU_ENGINE_AUTO || engine > NIPGPU_ENGINE_FACTOR) { delete m; return fail(NIPGPU_EINVAL, "unknown engine"); }
  m->engine = engine == NIPGPU_ENGINE_JTREE || engine == NIPGPU_ENGINE_FACTOR
                  ? engine
                  : (m->chain.ok ? NIPGPU_ENGINE_CHAIN : NIPGPU_ENGINE_JTREE);
  cudaDeviceProp prop;
  CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  m->sm_count = prop.multiProcessorCount;
  CREATE_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CREATE_CUDA(cudaEventCreate(&m->ev0));
  CREATE_CUDA(cudaEventCreate(&m->ev1));
  CREATE_CUDA(cudaEventCreate(&m->ev_mid));
  cudaStream_t st = m->stream;

  // ---- table offsets (int: the reference indexes tables with int too) ----
  if (hm.toff[hm.nc] > INT32_MAX) { nipgpu_model_destroy(m); return fail(NIPGPU_EUNSUPPORTED, "tables exceed 2^31 entries"); }
  m->tab_off.resize(hm.nc + 1);
  for (int c = 0; c <= hm.nc; c++) m->tab_off[c] = (int)hm.toff[c];

  // ---- projections -> one int pool ----
  std::vector<int> pool;
  std::vector<DProj> dp(hm.projs.size());
  for (size_t i = 0; i < hm.projs.size(); i++) {
    Proj& p = hm.projs[i];
    p.base_pos = (int)pool.size();
    pool.insert(pool.end(), p.base.begin(), p.base.end());
    p.off_pos = (int)pool.size();
    pool.insert(pool.end(), p.off.begin(), p.off.end());
    dp[i].tab = m->tab_off[p.clique]; dp[i].m = p.m; dp[i].R = p.R; dp[i].lanes = p.lanes;
    dp[i].base = p.base_pos; dp[i].off = p.off_pos;
    p.jlo_pos = (int)pool.size();
    pool.insert(pool.end(), p.jlo.begin(), p.jlo.end());
    p.jhi_pos = (int)pool.size();
    pool.insert(pool.end(), p.jhi.begin(), p.jhi.end());
    dp[i].clq = p.clique; dp[i].F = p.F; dp[i].jlo = p.jlo_pos; dp[i].jhi = p.jhi_pos;
  }
  auto to_dmsg = [&](const std::vector<Msg>& v) {
    std::vector<DMsg> r(v.size());
    for (size_t i = 0; i < v.size(); i++) r[i] = DMsg{v[i].proj_src, v[i].proj_dst, v[i].slot, v[i].size};
    return r;
  };
  int e = 0;
  std::vector<long long> coff(hm.coff.begin(), hm.coff.end());
  if ((e = dev_upload(&m->d_ipool, pool, st)) || (e = dev_upload(&m->d_projs, dp, st)) ||
      (e = dev_upload(&m->d_collect, to_dmsg(hm.collect), st)) ||
      (e = dev_upload(&m->d_distribute, to_dmsg(hm.distribute), st)) ||
      (e = dev_upload(&m->d_path, to_dmsg(hm.path_to_out), st)) ||
      (e = dev_upload(&m->d_proj_var, hm.proj_var, st)) || (e = dev_upload(&m->d_proj_fam, hm.proj_fam, st)) ||
      (e = dev_upload(&m->d_var_flags, hm.flags, st)) || (e = dev_upload(&m->d_coff, coff, st)) ||
      (e = dev_upload(&m->d_prior_off, hm.prior_off, st)) || (e = dev_upload(&m->d_prior_vars, hm.prior_vars, st)) ||
      (e = dev_upload(&m->d_orig, hm.tables, st)) || (e = dev_upload(&m->d_prior, hm.prior, st))) {
    nipgpu_model_destroy(m);
    return e;
  }
  CREATE_CUDA(cudaStreamSynchronize(st));  // the temporaries above die here
  const size_t T = (size_t)m->tab_off[hm.nc];
  CREATE_CUDA(cudaMalloc((void**)&m->d_prior_flags, std::max<size_t>(hm.prior_vars.size(), 1) * sizeof(int)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_base0, T * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_base1, T * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_R1, (size_t)hm.S * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_m10, sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_counts, (size_t)(hm.coff[hm.nv] + 2) * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_slice_start, T * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_slice_tab, T * sizeof(double)));
  CREATE_CUDA(cudaMalloc((void**)&m->d_slice_msg, (size_t)std::max(hm.msg_total, 1) * sizeof(double)));

  DProgram& P = m->prog;
  P.tab_total = (int)T; P.msg_total = hm.msg_total; P.msg_max = hm.msg_max;
  P.scratch = std::max({hm.fam_max, hm.card_max, hm.S, 1});
  P.n_collect = (int)hm.collect.size(); P.n_distribute = (int)hm.distribute.size();
  P.n_path = (int)hm.path_to_out.size();
  P.nif = hm.nif; P.S = hm.S; P.proj_in = hm.proj_in; P.proj_out = hm.proj_out;
  P.root_tab = m->tab_off[0]; P.root_size = hm.csize[0]; P.nv = hm.nv; P.n_cliques = hm.nc;
  P.projs = m->d_projs; P.ipool = m->d_ipool; P.collect = m->d_collect; P.distribute = m->d_distribute;
  P.path = m->d_path; P.base0 = m->d_base0; P.base1 = m->d_base1; P.R1 = m->d_R1; P.m1_0 = m->d_m10;
  P.proj_var = m->d_proj_var; P.proj_fam = m->d_proj_fam; P.coff = m->d_coff; P.var_flags = m->d_var_flags;

  if ((e = choose_launch(m)) || (e = chain_upload_structure(hm, m->chain, st))) {
    nipgpu_model_destroy(m);
    return e;
  }
  // engine 3 on request, or by itself for models whose clique tables do not fit shared memory
  // (engine 1 would keep them in per-CTA HBM workspaces or stream them with the whole grid;
  // measured, tools/dev_midsize.py: engine 3 is 5x faster at 6^6-entry cliques, 15x at 8^6 and
  // 12^6).  NIPGPU_FACTOR=0 keeps those models on engine 1.
  {
    const char* fenv = getenv("NIPGPU_FACTOR");
    const bool in_hbm = m->launch.mode == JT_MODE_GRID || (m->launch.mode == JT_MODE_CTA && m->launch.gwork != nullptr);
    const bool auto_fac = engine == NIPGPU_ENGINE_AUTO && !m->chain.ok && in_hbm && !(fenv && fenv[0] == '0');
    if (engine == NIPGPU_ENGINE_FACTOR || auto_fac) {
      fac_build(hm, m->fac);
      if (m->fac.ok) { m->engine = NIPGPU_ENGINE_FACTOR; m->fac_auto = engine == NIPGPU_ENGINE_AUTO; }
      else if (engine == NIPGPU_ENGINE_FACTOR) {
        const std::string why = m->fac.why;
        nipgpu_model_destroy(m);
        return fail(NIPGPU_EUNSUPPORTED, "factor engine: " + why);
      }
    }
  }
  if ((e = refresh_derived(m))) {
    nipgpu_model_destroy(m);
    return e;
  }
  m->lik.resize(hm.nv);
  for (int v = 0; v < hm.nv; v++) m->lik[v].assign(hm.card[v], 1.0);
  m->prior_entered.assign(hm.nv, 0);
  *out = m;
  return NIPGPU_OK;
}

void nipgpu_model_destroy(nipgpu_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (int g = 0; g < 8; g++) if (m->aux_stream[g]) cudaStreamDestroy(m->aux_stream[g]);
  for (int g = 0; g < 9; g++) if (m->aux_event[g]) cudaEventDestroy(m->aux_event[g]);
  cudaFree(m->d_trace); cudaFree(m->d_ipool); cudaFree(m->d_projs); cudaFree(m->d_collect); cudaFree(m->d_distribute);
  cudaFree(m->d_path); cudaFree(m->d_proj_var); cudaFree(m->d_proj_fam); cudaFree(m->d_var_flags);
  cudaFree(m->d_coff); cudaFree(m->d_prior_off); cudaFree(m->d_prior_vars); cudaFree(m->d_prior_flags);
  cudaFree(m->d_orig); cudaFree(m->d_prior); cudaFree(m->d_base0); cudaFree(m->d_base1);
  cudaFree(m->d_R1); cudaFree(m->d_m10); cudaFree(m->d_counts); cudaFree(m->d_acc); cudaFree(m->d_gwork);
  cudaFree(m->d_slice_start); cudaFree(m->d_slice_tab); cudaFree(m->d_slice_msg);
  if (m->prop_graph) cudaGraphExecDestroy(m->prop_graph);
  cudaFree(m->d_prop); cudaFree(m->d_vec);
  if (m->h_prop) cudaFreeHost(m->h_prop);
  chain_free(m->chain);
  fac_free(m->fac);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  for (int k = 0; k < 4; k++) if (m->chunk_ev[k]) cudaEventDestroy(m->chunk_ev[k]);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->ev_mid) cudaEventDestroy(m->ev_mid);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int nipgpu_model_engine(const nipgpu_model* m) { return m ? m->engine : 0; }

int nipgpu_model_factorable(const nipgpu_model_desc* desc) {
  HostModel hm;
  const std::string err = hm.load(desc);
  if (!err.empty()) return -fail(NIPGPU_EINVAL, "model description: " + err);
  FacEngine fe;
  fac_build(hm, fe);
  if (!fe.ok) set_error("factor engine: " + fe.why);
  return fe.ok ? 1 : 0;
}
void* nipgpu_model_stream(nipgpu_model* m) { return m ? (void*)m->stream : nullptr; }

int nipgpu_model_set_parameters(nipgpu_model* m, const double* tables, const double* prior) {
  if (!m || !tables) return fail(NIPGPU_EINVAL, "null argument");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  NIPGPU_CUDA(cudaMemcpyAsync(m->d_orig, tables, (size_t)m->prog.tab_total * sizeof(double),
                              cudaMemcpyHostToDevice, m->stream));
  if (prior && m->hm.prior_off[m->hm.nv] > 0)
    NIPGPU_CUDA(cudaMemcpyAsync(m->d_prior, prior, (size_t)m->hm.prior_off[m->hm.nv] * sizeof(double),
                                cudaMemcpyHostToDevice, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  if (m->fac.ok) {   // engine 3 works on the factors of the new tables
    std::copy(tables, tables + m->prog.tab_total, m->hm.tables.begin());
    fac_free(m->fac);
    fac_build(m->hm, m->fac);
    if (!m->fac.ok) {
      if (m->engine == NIPGPU_ENGINE_FACTOR) m->engine = NIPGPU_ENGINE_JTREE;
      set_error("factor engine dropped: " + m->fac.why);
    }
  }
  return refresh_derived(m);
}

int nipgpu_model_get_parameters(nipgpu_model* m, double* tables, double* prior) {
  if (!m) return fail(NIPGPU_EINVAL, "null argument");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  if (tables)
    NIPGPU_CUDA(cudaMemcpyAsync(tables, m->d_orig, (size_t)m->prog.tab_total * sizeof(double),
                                cudaMemcpyDeviceToHost, m->stream));
  if (prior && m->hm.prior_off[m->hm.nv] > 0)
    NIPGPU_CUDA(cudaMemcpyAsync(prior, m->d_prior, (size_t)m->hm.prior_off[m->hm.nv] * sizeof(double),
                                cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  return NIPGPU_OK;
}

int nipgpu_batch_create(nipgpu_model* m, int32_t n_series, const int32_t* lengths, int32_t n_obs,
                        const int32_t* obs_vars, const int32_t* data, nipgpu_batch** out) {
  if (!m || !out || n_series < 0 || n_obs < 0) return fail(NIPGPU_EINVAL, "bad batch arguments");
  *out = nullptr;
  NIPGPU_CUDA(cudaSetDevice(m->device));
  nipgpu_batch* b = new nipgpu_batch();
  b->m = m; b->n_series = n_series; b->n_obs = n_obs;
  b->len.assign(lengths, lengths + n_series);
  b->obs_vars.assign(obs_vars, obs_vars + n_obs);
  for (int k = 0; k < n_obs; k++)
    if (obs_vars[k] < 0 || obs_vars[k] >= m->hm.nv) { delete b; return fail(NIPGPU_EINVAL, "observed variable out of range"); }
  b->row_off.resize(n_series + 1);
  b->rows = 0; b->t_max = 0;
  for (int s = 0; s < n_series; s++) {
    if (lengths[s] < 0) { delete b; return fail(NIPGPU_EINVAL, "negative series length"); }
    b->row_off[s] = b->rows;
    b->rows += lengths[s];
    b->t_max = std::max(b->t_max, lengths[s]);
  }
  b->row_off[n_series] = b->rows;
  // range check of the observations (the reference would index out of bounds)
  for (long long r = 0; r < b->rows; r++)
    for (int k = 0; k < n_obs; k++)
      if (data[r * n_obs + k] >= m->hm.card[obs_vars[k]]) { delete b; return fail(NIPGPU_EINVAL, "observation index >= cardinality"); }
  cudaStream_t st = m->stream;
  int e = 0;
  if ((e = dev_upload(&b->d_len, b->len, st)) || (e = dev_upload(&b->d_row_off, b->row_off, st)) ||
      (e = dev_upload(&b->d_obs, data, (size_t)b->rows * n_obs, st))) {
    nipgpu_batch_destroy(b);
    return e;
  }
  NIPGPU_CUDA(cudaMalloc((void**)&b->d_obs_proj, 3 * (size_t)std::max(n_obs, 1) * sizeof(int)));
  NIPGPU_CUDA(cudaMalloc((void**)&b->d_ll, (size_t)std::max(n_series, 1) * sizeof(double)));
  NIPGPU_CUDA(cudaMalloc((void**)&b->d_status, (size_t)std::max(n_series, 1) * sizeof(int)));
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  *out = b;
  return NIPGPU_OK;
}

int nipgpu_batch_update(nipgpu_batch* b, const int32_t* data) {
  if (!b || !data) return fail(NIPGPU_EINVAL, "null argument");
  NIPGPU_CUDA(cudaSetDevice(b->m->device));
  cudaStream_t st = b->m->stream;
  NIPGPU_CUDA(cudaMemcpyAsync(b->d_obs, data, (size_t)b->rows * b->n_obs * sizeof(int), cudaMemcpyHostToDevice, st));
  b->chain.plan_key.clear();  // cached per-row evidence configuration is stale
  b->chain.cfgT_key.clear();  // ... and its time-major copy
  // range check on the device (an index >= the cardinality would address past the evidence
  // tables; the reference would index out of bounds, src/nip.c:994)
  if (!b->d_check) {
    std::vector<int> cc(std::max(b->n_obs, 1) + 1, 0);
    for (int k = 0; k < b->n_obs; k++) cc[k + 1] = b->m->hm.card[b->obs_vars[k]];
    if (int e = dev_upload(&b->d_check, cc, st)) return e;   // [0] flag, [1..] cardinality per column
    NIPGPU_CUDA(cudaStreamSynchronize(st));
  }
  int flag = 0;
  NIPGPU_CUDA(cudaMemsetAsync(b->d_check, 0, sizeof(int), st));
  if (int e = check_obs(b->d_obs, b->rows, b->n_obs, b->d_check + 1, b->d_check, st)) return e;
  NIPGPU_CUDA(cudaMemcpyAsync(&flag, b->d_check, sizeof(int), cudaMemcpyDeviceToHost, st));
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  if (flag) return fail(NIPGPU_EINVAL, "observation index >= cardinality");
  return NIPGPU_OK;
}

void nipgpu_batch_destroy(nipgpu_batch* b) {
  if (!b) return;
  if (b->m) { cudaSetDevice(b->m->device); cudaStreamSynchronize(b->m->stream); }
  cudaFree(b->d_len); cudaFree(b->d_row_off); cudaFree(b->d_obs); cudaFree(b->d_obs_proj);
  cudaFree(b->d_qproj); cudaFree(b->d_qoff); cudaFree(b->d_alpha); cudaFree(b->d_post);
  cudaFree(b->d_ll); cudaFree(b->d_like); cudaFree(b->d_status); cudaFree(b->d_first); cudaFree(b->d_joint); cudaFree(b->d_check);
  chain_batch_free(b->chain);
  delete b;
}

int nipgpu_infer_device(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int32_t n_query,
                        const int32_t* query_vars, int forward_only, int want_loglik,
                        double** post_dev, double** loglik_dev) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  if (int e = infer_impl(m, b, use_evidence, n_query, query_vars, forward_only, want_loglik, true)) return e;
  if (post_dev) *post_dev = b->d_post;
  if (loglik_dev) *loglik_dev = b->d_ll;
  return NIPGPU_OK;
}

int nipgpu_infer(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int32_t n_query,
                 const int32_t* query_vars, int forward_only, double* post, double* loglik) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  bool copied = false;
  if (int e = infer_impl(m, b, use_evidence, n_query, query_vars, forward_only, loglik != nullptr, post != nullptr,
                         post, &copied)) return e;
  if (post && n_query > 0 && !copied) {
    size_t row = 0;
    for (int i = 0; i < n_query; i++) row += m->hm.card[query_vars[i]];
    NIPGPU_CUDA(cudaMemcpyAsync(post, b->d_post, (size_t)b->rows * row * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  }
  if (loglik)
    NIPGPU_CUDA(cudaMemcpyAsync(loglik, b->d_ll, (size_t)b->n_series * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  return NIPGPU_OK;
}

int64_t nipgpu_model_counts_size(const nipgpu_model* m) { return m ? m->hm.coff[m->hm.nv] : 0; }

int nipgpu_model_counts_offsets(const nipgpu_model* m, int64_t* off) {
  if (!m || !off) return fail(NIPGPU_EINVAL, "null argument");
  for (int v = 0; v <= m->hm.nv; v++) off[v] = m->hm.coff[v];
  return NIPGPU_OK;
}

int nipgpu_em_counts_device(nipgpu_model* m, double** counts_dev, int64_t* n_doubles) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  if (counts_dev) *counts_dev = m->d_counts;
  if (n_doubles) *n_doubles = m->hm.coff[m->hm.nv] + 2;
  return NIPGPU_OK;
}

}  // extern "C"

namespace nipgpu {
// E-step of one model on its own stream, nothing awaited: when this returns the kernels are
// enqueued and m->d_counts (+ loglik, status) will hold this device's share.
int estep_enqueue(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int add_pseudocount) {
  if (!m || !b || b->m != m) return fail(NIPGPU_EINVAL, "model/batch mismatch");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const HostModel& hm = m->hm;
  const long long n = hm.coff[hm.nv];
  // ---- engine 2: counts as DMMA GEMMs ----
  ChainPlan plan;
  if (m->engine == NIPGPU_ENGINE_CHAIN && m->chain.ok &&
      chain_plan(hm, m->chain, b->n_obs, b->obs_vars.data(), use_evidence, plan)) {
    if (int e = chain_batch_prepare(m->chain, b->chain, b->n_series, b->len.data(), b->rows, b->t_max, m->stream)) return e;
    ChainEmArgs x;
    x.base.n_series = b->n_series; x.base.n_obs = b->n_obs; x.base.t_max = b->t_max; x.base.rows = b->rows;
    x.base.d_obs = b->d_obs; x.base.d_row_off = b->d_row_off; x.base.want_ll = 1; x.base.forward_only = 0;
    x.base.d_post = nullptr; x.base.post_stride = 0; x.base.post_off = 0;
    x.base.d_ll = b->d_ll; x.base.d_status = b->d_status;
    x.d_base0 = m->d_base0; x.d_base1 = m->d_base1; x.tab_off = &m->tab_off; x.d_ipool = m->d_ipool;
    x.pseudo = add_pseudocount ? 1.0 : 0.0; x.d_counts = m->d_counts; x.sm_count = m->sm_count;
    const int rc = chain_estep(hm, m->chain, b->chain, plan, x, m->stream, m->ev0, m->ev1);
    if (rc == NIPGPU_OK) { m->last_kernel_n = 7; return NIPGPU_OK; }
    if (rc != NIPGPU_EUNSUPPORTED) return rc;
  }
  if (m->engine == NIPGPU_ENGINE_FACTOR && m->fac.ok) {
    const int slots = fac_slots(hm, m->fac, b->n_series);
    if (m->acc_groups < (size_t)slots) {
      cudaFree(m->d_acc);
      m->d_acc = nullptr;
      m->acc_groups = 0;
      NIPGPU_CUDA(cudaMalloc((void**)&m->d_acc, (size_t)slots * (size_t)n * sizeof(double)));
      m->acc_groups = slots;
    }
    NIPGPU_CUDA(cudaMemsetAsync(m->d_acc, 0, (size_t)slots * n * sizeof(double), m->stream));
    FacRunArgs a{};
    fill_fac_args(m, b, use_evidence, &a);
    a.want_ll = 1; a.forward_only = 0; a.d_acc = m->d_acc; a.acc_stride = n; a.acc_slots = slots;
    NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
    const int fe = fac_run(hm, m->fac, a, m->stream);
    if (fe == NIPGPU_EUNSUPPORTED && m->fac_auto) {
      m->engine = NIPGPU_ENGINE_JTREE;   // see infer_impl
    } else {
      if (fe) return fe;
      NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
      if (int e = finish_estep(m->d_acc, slots, n, n, add_pseudocount ? 1.0 : 0.0, b->d_ll, b->d_status,
                               b->n_series, m->d_counts, m->stream)) return e;
      m->last_kernel_n = 0;
      return NIPGPU_OK;
    }
  }
  const int* obs_proj = nullptr;
  if (int e = upload_obs_proj(m, b, use_evidence, true, 0, &obs_proj)) return e;
  if (int e = ensure_alpha(m, b)) return e;
  const DBatch B = dev_batch(b, obs_proj);
  const JtLaunch l = jt_fit(m->launch, b->n_series);
  if (m->acc_groups < (size_t)l.slots) {
    cudaFree(m->d_acc);
    m->acc_groups = l.slots;
    NIPGPU_CUDA(cudaMalloc((void**)&m->d_acc, m->acc_groups * (size_t)n * sizeof(double)));
  }
  NIPGPU_CUDA(cudaMemsetAsync(m->d_acc, 0, (size_t)l.slots * n * sizeof(double), m->stream));
  DQuery Q{};
  NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
  if (int e = jt_forward(m->prog, B, Q, l, 1, 0, b->d_alpha, nullptr, b->d_ll, b->d_status, m->stream)) return e;
  if (int e = jt_backward(m->prog, B, Q, l, b->d_alpha, nullptr, m->d_acc, n, m->stream)) return e;
  NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
  if (int e = finish_estep(m->d_acc, l.slots, n, n, add_pseudocount ? 1.0 : 0.0, b->d_ll, b->d_status,
                           b->n_series, m->d_counts, m->stream)) return e;
  m->last_kernel_n = 2;
  return NIPGPU_OK;
}

// waits for the stream and brings the scalars (and, on request, the counts) home
int estep_finish(nipgpu_model* m, double* counts, double* loglik, int* status) {
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const long long n = m->hm.coff[m->hm.nv];
  double tail[2] = {0, 0};
  NIPGPU_CUDA(cudaMemcpyAsync(tail, m->d_counts + n, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  if (counts)
    NIPGPU_CUDA(cudaMemcpyAsync(counts, m->d_counts, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_kernel_ms = ms;
  if (loglik) *loglik = tail[0];
  if (status) *status = tail[1] != 0 ? NIPGPU_EBADLUCK : 0;
  return NIPGPU_OK;
}
}  // namespace nipgpu

extern "C" {

int nipgpu_em_estep(nipgpu_model* m, nipgpu_batch* b, const uint8_t* use_evidence, int add_pseudocount,
                    double* counts, double* loglik, int* status) {
  if (int e = estep_enqueue(m, b, use_evidence, add_pseudocount)) return e;
  return estep_finish(m, counts, loglik, status);
}

int nipgpu_em_mstep(nipgpu_model* m, const double* counts) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const HostModel& hm = m->hm;
  cudaStream_t st = m->stream;
  const long long n = hm.coff[hm.nv];
  if (counts) {
    NIPGPU_CUDA(cudaMemcpyAsync(m->d_counts, counts, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    NIPGPU_CUDA(cudaStreamSynchronize(st));
  }
  // 1. normalise_cpd over the child dimension (src/nip.c:2032-2038)
  for (int v = 0; v < hm.nv; v++)
    if (int e = normalise_cpd(m->d_counts + hm.coff[v], hm.coff[v + 1] - hm.coff[v], hm.card[v], st)) return e;
  // 2. total_reset: every original_p <- 1 (src/nip.c:76-85)
  if (int e = fill(m->d_orig, m->prog.tab_total, 1.0, st)) return e;
  // 3. CPTs into family cliques, priors of parentless variables (src/nip.c:2044-2067)
  for (int v = 0; v < hm.nv; v++) {
    if (hm.nparents(v) > 0) {
      const int c = hm.family[v];
      FamMap fm;
      fm.n = 1 + hm.nparents(v);
      int fstride = 1;
      for (int k = 0; k < fm.n; k++) {
        const int var = k == 0 ? v : hm.parents[hm.poff[v] + k - 1];
        int cs = 1;
        for (int j = 0; j < hm.var_pos(c, var); j++) cs *= hm.card[hm.clique_vars(c)[j]];
        fm.cstride[k] = cs; fm.card[k] = hm.card[var]; fm.fstride[k] = fstride;
        fstride *= hm.card[var];
      }
      if (int e = init_potential(m->d_orig + m->tab_off[c], hm.csize[c], m->d_counts + hm.coff[v], fm, st)) return e;
    } else {
      NIPGPU_CUDA(cudaMemcpyAsync(m->d_prior + hm.prior_off[v], m->d_counts + hm.coff[v],
                                  (size_t)hm.card[v] * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
  }
  return refresh_derived(m, true);
}

int nipgpu_likelihood(nipgpu_model* m, nipgpu_batch* b, const uint8_t* evidence_off,
                      const uint8_t* evidence_on, double* out) {
  if (!m || !b || b->m != m || !out) return fail(NIPGPU_EINVAL, "bad arguments");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const int *p_off = nullptr, *p_on = nullptr;
  if (int e = upload_obs_proj(m, b, evidence_off, false, 1, &p_off)) return e;
  if (int e = upload_obs_proj(m, b, evidence_on, false, 2, &p_on)) return e;
  if (!b->d_like) NIPGPU_CUDA(cudaMalloc((void**)&b->d_like, std::max<size_t>((size_t)b->rows * 2, 1) * sizeof(double)));
  const DBatch B = dev_batch(b, nullptr);
  const JtLaunch l = jt_fit(m->launch, b->n_series);

  // ---- few evidence configurations: evaluate each once, the records gather ----
  const HostModel& hm = m->hm;
  std::vector<int> col_stride(std::max(b->n_obs, 1), 0), col_card(std::max(b->n_obs, 1), 1);
  long long n_cfg = 1;
  for (int k = 0; k < b->n_obs && n_cfg <= (1 << 16); k++) {
    const int v = b->obs_vars[k];
    if ((evidence_off && evidence_off[v]) || (evidence_on && evidence_on[v])) {
      col_stride[k] = (int)n_cfg;
      col_card[k] = hm.card[v];
      n_cfg *= hm.card[v] + 2;  // missing, every state, out of range
    }
  }
  const char* no_memo = getenv("NIPGPU_NO_LIKELIHOOD_MEMO");
  if (n_cfg <= (1 << 16) && b->rows >= 4 * n_cfg && !(no_memo && no_memo[0] == '1')) {
    const int nc = (int)n_cfg, no = std::max(b->n_obs, 1);
    std::vector<int> len(nc, 2), obs((size_t)nc * 2 * no, -1);
    std::vector<long long> off(nc);
    for (int c = 0; c < nc; c++) {
      off[c] = 2LL * c;
      for (int k = 0; k < b->n_obs; k++) {
        if (!col_stride[k]) continue;
        const int d = (c / col_stride[k]) % (col_card[k] + 2);  // 0 missing, 1..card states, card+1 out of range
        const int o = d == 0 ? -1 : (d <= col_card[k] ? d - 1 : col_card[k]);
        obs[(size_t)(2 * c) * no + k] = o;
        obs[(size_t)(2 * c + 1) * no + k] = o;
      }
    }
    int *d_len = nullptr, *d_obs = nullptr, *d_cs = nullptr, *d_cc = nullptr;
    long long* d_off = nullptr;
    double* d_table = nullptr;
    auto cleanup = [&]() { cudaFree(d_len); cudaFree(d_obs); cudaFree(d_cs); cudaFree(d_cc); cudaFree(d_off); cudaFree(d_table); };
    int e = 0;
    if ((e = dev_upload(&d_len, len, m->stream)) || (e = dev_upload(&d_obs, obs, m->stream)) ||
        (e = dev_upload(&d_off, off, m->stream)) || (e = dev_upload(&d_cs, col_stride, m->stream)) ||
        (e = dev_upload(&d_cc, col_card, m->stream))) { cleanup(); return e; }
    if (cudaMalloc((void**)&d_table, (size_t)nc * 4 * sizeof(double)) != cudaSuccess) { cleanup(); return fail(NIPGPU_ENOMEM, "likelihood table"); }
    if (!b->d_first) {
      NIPGPU_CUDA(cudaMalloc((void**)&b->d_first, std::max<long long>(b->rows, 1)));
      NIPGPU_CUDA(cudaMemsetAsync(b->d_first, 0, std::max<long long>(b->rows, 1), m->stream));
      if ((e = jt_first_rows(b->d_row_off, b->n_series, b->rows, b->d_first, m->stream))) { cleanup(); return e; }
    }
    DBatch S;
    S.n_series = nc; S.n_obs = b->n_obs; S.len = d_len; S.row_off = d_off; S.obs = d_obs; S.obs_proj = nullptr;
    NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
    if ((e = jt_likelihood(m->prog, S, p_off, p_on, jt_fit(m->launch, nc), d_table, m->stream)) ||
        (e = jt_like_gather(b->d_obs, b->n_obs, b->rows, b->d_first, d_cs, d_cc, d_table, b->d_like, m->sm_count, m->stream,
                            col_stride[0], col_card[0]))) {
      cleanup();
      return e;
    }
    NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
    NIPGPU_CUDA(cudaMemcpyAsync(out, b->d_like, (size_t)b->rows * 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
    cleanup();
    float ms = 0;
    if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) { m->last_kernel_ms = ms; m->last_kernel_n = 2; }
    return NIPGPU_OK;
  }

  NIPGPU_CUDA(cudaEventRecord(m->ev0, m->stream));
  if (int e = jt_likelihood(m->prog, B, p_off, p_on, l, b->d_like, m->stream)) return e;
  NIPGPU_CUDA(cudaEventRecord(m->ev1, m->stream));
  NIPGPU_CUDA(cudaMemcpyAsync(out, b->d_like, (size_t)b->rows * 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) { m->last_kernel_ms = ms; m->last_kernel_n = 1; }
  return NIPGPU_OK;
}

// ---- single-slice stateful API ------------------------------------------
int nipgpu_slice_reset(nipgpu_model* m) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  for (auto& l : m->lik) std::fill(l.begin(), l.end(), 1.0);
  std::fill(m->prior_entered.begin(), m->prior_entered.end(), 0);
  m->slice_consistent = false;
  return NIPGPU_OK;
}

int nipgpu_slice_use_priors(nipgpu_model* m, int has_history) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  for (int v : m->hm.prior_vars)
    if (!m->prior_entered[v] && (!has_history || !(m->hm.flags[v] & NIPGPU_IF_OLD_OUTGOING)))
      m->prior_entered[v] = 1;
  m->slice_consistent = false;
  return NIPGPU_OK;
}

int nipgpu_slice_enter_prior(nipgpu_model* m, int32_t var) {
  if (!m || var < 0 || var >= m->hm.nv || m->hm.nparents(var) != 0) return fail(NIPGPU_EINVAL, "bad arguments");
  m->prior_entered[var] = 1;
  m->slice_consistent = false;
  return NIPGPU_OK;
}

int nipgpu_slice_get_sepset(nipgpu_model* m, int32_t sepset, double* out) {
  if (!m || sepset < 0 || sepset >= m->hm.ns || !out) return fail(NIPGPU_EINVAL, "bad arguments");
  if (!m->slice_consistent) return fail(NIPGPU_EINVAL, "call nipgpu_slice_make_consistent first");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  NIPGPU_CUDA(cudaMemcpy(out, m->d_slice_msg + m->hm.sep_slot[sepset], (size_t)m->hm.ssize[sepset] * sizeof(double), cudaMemcpyDeviceToHost));
  return NIPGPU_OK;
}

int nipgpu_slice_enter_evidence(nipgpu_model* m, int32_t var, const double* likelihood) {
  if (!m || var < 0 || var >= m->hm.nv || !likelihood) return fail(NIPGPU_EINVAL, "bad arguments");
  // nip_enter_evidence replaces the old likelihood by the new one (old is divided out,
  // or everything is retracted and re-entered): the net effect is "current = new".
  m->lik[var].assign(likelihood, likelihood + m->hm.card[var]);
  m->slice_consistent = false;
  return NIPGPU_OK;
}

int nipgpu_slice_make_consistent(nipgpu_model* m) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const HostModel& hm = m->hm;
  cudaStream_t st = m->stream;
  NIPGPU_CUDA(cudaMemcpyAsync(m->d_slice_start, m->d_orig, (size_t)m->prog.tab_total * sizeof(double),
                              cudaMemcpyDeviceToDevice, st));
  auto stride_of = [&](int c, int v) {
    int s = 1;
    for (int j = 0; j < hm.var_pos(c, v); j++) s *= hm.card[hm.clique_vars(c)[j]];
    return s;
  };
  for (size_t k = 0; k < hm.prior_vars.size(); k++) {
    const int v = hm.prior_vars[k], c = hm.family[v];
    if (!m->prior_entered[v]) continue;
    if (int e = apply_vector(m->d_slice_start + m->tab_off[c], hm.csize[c], stride_of(c, v), hm.card[v],
                             m->d_prior + hm.prior_off[v], m->d_prior_flags + k, st)) return e;
  }
  // all evidence vectors in one upload (persistent buffer, one slot of card_max per variable)
  if (!m->d_vec) NIPGPU_CUDA(cudaMalloc((void**)&m->d_vec, (size_t)hm.nv * hm.card_max * sizeof(double)));
  m->lik_stage.assign((size_t)hm.nv * hm.card_max, 1.0);
  bool any = false;
  for (int v = 0; v < hm.nv; v++) {
    bool all_one = true;
    for (double x : m->lik[v]) all_one = all_one && x == 1.0;
    if (all_one) continue;
    any = true;
    std::copy(m->lik[v].begin(), m->lik[v].end(), m->lik_stage.begin() + (size_t)v * hm.card_max);
  }
  if (any) {
    NIPGPU_CUDA(cudaMemcpyAsync(m->d_vec, m->lik_stage.data(), m->lik_stage.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int v = 0; v < hm.nv; v++) {
      bool all_one = true;
      for (double x : m->lik[v]) all_one = all_one && x == 1.0;
      if (all_one) continue;
      const int c = hm.family[v];
      if (int e = apply_vector(m->d_slice_start + m->tab_off[c], hm.csize[c], stride_of(c, v), hm.card[v],
                               m->d_vec + (size_t)v * hm.card_max, nullptr, st)) return e;
    }
  }
  if (int e = jt_slice(m->prog, m->launch, m->d_slice_start, m->d_slice_tab, m->d_slice_msg, st)) return e;
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  m->slice_consistent = true;
  return NIPGPU_OK;
}

// make_consistent (src/nip.c:1600-1617) as one device transaction on the caller's own tree
// state.  First call: buffers + a captured graph; every call: memcpy into the pinned block, one
// graph launch (H2D, k_jt_propagate, D2H), one sync, memcpy out.
int nipgpu_slice_propagate(nipgpu_model* m, const double* clique_tables, const double* sepset_tables,
                           double* clique_out, double* sepset_new_out, double* sepset_old_out) {
  if (!m || !clique_tables || !clique_out) return fail(NIPGPU_EINVAL, "bad arguments");
  const size_t T = (size_t)m->prog.tab_total, S = (size_t)m->hm.msg_total;
  if (S > 0 && !sepset_tables) return fail(NIPGPU_EINVAL, "bad arguments");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  const size_t n_in = T + S, n_out = T + 2 * S;
  cudaStream_t st = m->stream;
  if (!m->h_prop) {
    NIPGPU_CUDA(cudaMallocHost((void**)&m->h_prop, (n_in + n_out) * sizeof(double)));
    NIPGPU_CUDA(cudaMalloc((void**)&m->d_prop, (n_in + n_out) * sizeof(double)));
  }
  memcpy(m->h_prop, clique_tables, T * sizeof(double));
  if (S) memcpy(m->h_prop + T, sepset_tables, S * sizeof(double));
  auto enqueue = [&]() -> int {
    NIPGPU_CUDA(cudaMemcpyAsync(m->d_prop, m->h_prop, n_in * sizeof(double), cudaMemcpyHostToDevice, st));
    if (int e = jt_propagate(m->prog, m->launch, m->d_prop, m->d_prop + n_in, st)) return e;
    NIPGPU_CUDA(cudaMemcpyAsync(m->h_prop + n_in, m->d_prop + n_in, n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    return NIPGPU_OK;
  };
  if (m->launch.mode == JT_MODE_GRID) {   // cooperative launches are not captured
    if (int e = enqueue()) return e;
  } else {
    if (!m->prop_graph) {
      cudaGraph_t g = nullptr;
      NIPGPU_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int e = enqueue();
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (e) { if (g) cudaGraphDestroy(g); return e; }
      NIPGPU_CUDA(ce);
      const cudaError_t ie = cudaGraphInstantiate(&m->prop_graph, g, 0);
      cudaGraphDestroy(g);
      NIPGPU_CUDA(ie);
    } else {
      g_launches++;   // the kernel inside the graph
    }
    NIPGPU_CUDA(cudaGraphLaunch(m->prop_graph, st));
  }
  NIPGPU_CUDA(cudaStreamSynchronize(st));
  const double* o = m->h_prop + n_in;
  memcpy(clique_out, o, T * sizeof(double));
  if (S && sepset_new_out) memcpy(sepset_new_out, o + T, S * sizeof(double));
  if (S && sepset_old_out) memcpy(sepset_old_out, o + T + S, S * sizeof(double));
  return NIPGPU_OK;
}

int nipgpu_slice_get_clique(nipgpu_model* m, int32_t clique, double* out) {
  if (!m || clique < 0 || clique >= m->hm.nc || !out) return fail(NIPGPU_EINVAL, "bad arguments");
  if (!m->slice_consistent) return fail(NIPGPU_EINVAL, "call nipgpu_slice_make_consistent first");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  NIPGPU_CUDA(cudaMemcpy(out, m->d_slice_tab + m->tab_off[clique], (size_t)m->hm.csize[clique] * sizeof(double), cudaMemcpyDeviceToHost));
  return NIPGPU_OK;
}

int nipgpu_slice_mass(nipgpu_model* m, double* mass) {
  if (!m || !mass) return fail(NIPGPU_EINVAL, "bad arguments");
  if (!m->slice_consistent) return fail(NIPGPU_EINVAL, "call nipgpu_slice_make_consistent first");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  double* d_out = m->d_slice_start;  // scratch: the start tables are dead once the slice is consistent
  if (int e = jt_mass(m->d_slice_tab, m->prog.tab_total, m->d_slice_msg, m->hm.msg_total, d_out, m->stream)) return e;
  NIPGPU_CUDA(cudaMemcpyAsync(mass, d_out, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  return NIPGPU_OK;
}

int nipgpu_slice_marginal(nipgpu_model* m, int32_t var, double* out) {
  if (!m || var < 0 || var >= m->hm.nv || !out) return fail(NIPGPU_EINVAL, "bad arguments");
  if (!m->slice_consistent) return fail(NIPGPU_EINVAL, "call nipgpu_slice_make_consistent first");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  double* d_out = m->d_slice_start;
  if (int e = jt_marginal(m->prog, m->d_slice_tab, m->hm.proj_var[var], d_out, m->stream)) return e;
  NIPGPU_CUDA(cudaMemcpyAsync(out, d_out, (size_t)m->hm.card[var] * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  return NIPGPU_OK;
}

int nipgpu_sample(nipgpu_model* m, int32_t n_series, int32_t length, uint64_t seed, int32_t* out) {
  if (!m || !out || n_series < 0 || length < 0) return fail(NIPGPU_EINVAL, "bad arguments");
  NIPGPU_CUDA(cudaSetDevice(m->device));
  if (!m->chain.ok) return fail(NIPGPU_EUNSUPPORTED, "sampling needs a chain-structured model");
  const size_t n = (size_t)n_series * length * m->hm.nv;
  if (n == 0) return NIPGPU_OK;
  int* d_out = nullptr;
  NIPGPU_CUDA(cudaMalloc((void**)&d_out, n * sizeof(int)));
  int e = chain_sample(m->hm, m->chain, m->d_base0, m->d_base1, m->tab_off, m->d_ipool, n_series, length, seed,
                       d_out, m->stream);
  if (e == NIPGPU_OK && cudaMemcpy(out, d_out, n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) e = NIPGPU_ECUDA;
  cudaFree(d_out);
  if (e == NIPGPU_EUNSUPPORTED) return fail(e, "sampling: model layout not supported");
  return e;
}

int nipgpu_jt_trace(nipgpu_model* m, uint64_t* out, int cap_records, int reset) {
  if (!m || !out || cap_records < 0) return fail(NIPGPU_EINVAL, "bad arguments") , -1;
  if (!m->d_trace) return 0;
  NIPGPU_CUDA(cudaSetDevice(m->device));
  NIPGPU_CUDA(cudaStreamSynchronize(m->stream));
  std::vector<unsigned long long> h(JT_TRACE_WORDS);
  NIPGPU_CUDA(cudaMemcpy(h.data(), m->d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  const int n = (int)std::min<unsigned long long>(h[0], (unsigned long long)cap_records);
  for (int i = 0; i < 2 * n; i++) out[i] = h[1 + i];
  if (reset) NIPGPU_CUDA(cudaMemset(m->d_trace, 0, sizeof(unsigned long long)));
  return n;
}

int64_t nipgpu_launch_count(int reset) {
  const int64_t n = g_launches.load();
  if (reset) g_launches.store(0);
  return n;
}

int nipgpu_last_forward_ms(nipgpu_model* m, double* ms) {
  if (!m || !ms) return fail(NIPGPU_EINVAL, "bad arguments");
  *ms = m->last_forward_ms;
  return NIPGPU_OK;
}

int nipgpu_last_kernel_ms(nipgpu_model* m, double* ms, int32_t* n) {
  if (!m) return fail(NIPGPU_EINVAL, "null model");
  if (ms) *ms = m->last_kernel_ms;
  if (n) *n = m->last_kernel_n;
  return NIPGPU_OK;
}

}  // extern "C"
