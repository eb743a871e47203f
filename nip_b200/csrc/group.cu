// group.cu — the same model on several B200s of one box: sequences are sharded
// by the caller, every device runs the E-step of its shard, and the expected-count accumulators
// (+ log-likelihood, status) are summed by ONE ncclAllReduce over NVLink per EM iteration; the
// M-step then runs redundantly on every device so parameters never leave HBM.
//
// One host thread per device for the duration of a call (launch + the M-step's own
// synchronisation would otherwise serialise over the devices).  NCCL is bound at run time
// (dlopen of libnccl.so.2): single-device users of libnipgpu.so never load it, and a process
// that already carries an NCCL (torch) shares that one.
#include <dlfcn.h>
#include <nccl.h>

#include <thread>
#include <vector>

#include "api.cuh"

namespace nipgpu {
namespace {

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !AllReduce || !GetErrorString) { why = "libnccl lacks a symbol"; lib = nullptr; return false; }
    return true;
  }
};
Nccl g_nccl;

}  // namespace
}  // namespace nipgpu

using namespace nipgpu;

struct nipgpu_group {
  std::vector<nipgpu_model*> models;
  std::vector<ncclComm_t> comms;
};

namespace {
int gfail(int code, const std::string& msg) { set_error(msg); return code; }

// runs f(i) for every member on its own thread; the first non-zero code wins, its message is
// carried back to the caller's thread (nipgpu_last_error is thread-local)
template <class F>
int for_each_device(nipgpu_group* g, F f) {
  const int n = (int)g->models.size();
  std::vector<int> rc(n, 0);
  std::vector<std::string> msg(n);
  auto body = [&](int i) {
    rc[i] = f(i);
    if (rc[i]) msg[i] = nipgpu_last_error();
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++) th.emplace_back(body, i);
  body(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i]) return gfail(rc[i], "device " + std::to_string(g->models[i]->device) + ": " + msg[i]);
  return NIPGPU_OK;
}
}  // namespace

extern "C" {

int nipgpu_group_create(nipgpu_model** models, int n, nipgpu_group** out) {
  if (!models || n < 1 || !out) return gfail(NIPGPU_EINVAL, "bad arguments");
  *out = nullptr;
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) {
    if (!models[i]) return gfail(NIPGPU_EINVAL, "null model in group");
    devs[i] = models[i]->device;
    if (models[i]->hm.coff[models[i]->hm.nv] != models[0]->hm.coff[models[0]->hm.nv] ||
        models[i]->prog.tab_total != models[0]->prog.tab_total)
      return gfail(NIPGPU_EINVAL, "group members were compiled from different models");
    for (int j = 0; j < i; j++)
      if (devs[j] == devs[i]) return gfail(NIPGPU_EINVAL, "two group members on one device");
  }
  nipgpu_group* g = new nipgpu_group();
  g->models.assign(models, models + n);
  if (n > 1) {
    if (!g_nccl.load()) { delete g; return gfail(NIPGPU_EUNSUPPORTED, "NCCL not available: " + g_nccl.why); }
    g->comms.resize(n);
    const ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), n, devs.data());
    if (r != ncclSuccess) { delete g; return gfail(NIPGPU_ECUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); }
  }
  *out = g;
  return NIPGPU_OK;
}

void nipgpu_group_destroy(nipgpu_group* g) {
  if (!g) return;
  for (size_t i = 0; i < g->comms.size(); i++) {
    cudaSetDevice(g->models[i]->device);
    cudaStreamSynchronize(g->models[i]->stream);
    g_nccl.CommDestroy(g->comms[i]);
  }
  delete g;
}

int nipgpu_group_size(const nipgpu_group* g) { return g ? (int)g->models.size() : 0; }

int nipgpu_group_em_estep(nipgpu_group* g, nipgpu_batch** batches, const uint8_t* use_evidence,
                          int add_pseudocount, double* counts, double* loglik, int* status) {
  if (!g || !batches) return gfail(NIPGPU_EINVAL, "bad arguments");
  const int n = (int)g->models.size();
  const size_t len = (size_t)g->models[0]->hm.coff[g->models[0]->hm.nv] + 2;
  const int rc = for_each_device(g, [&](int i) -> int {
    nipgpu_model* m = g->models[i];
    // the reference's 1.0 pseudo-count (src/nip.c:2171-2172) enters the sum exactly once
    if (int e = estep_enqueue(m, batches[i], use_evidence, add_pseudocount && i == 0)) return e;
    if (n > 1) {
      const ncclResult_t r = g_nccl.AllReduce(m->d_counts, m->d_counts, len, ncclDouble, ncclSum, g->comms[i], m->stream);
      if (r != ncclSuccess) return gfail(NIPGPU_ECUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
    }
    // every device waits for its own stream (the all-reduce included); member 0 reports
    return estep_finish(m, i == 0 ? counts : nullptr, i == 0 ? loglik : nullptr, i == 0 ? status : nullptr);
  });
  return rc;
}

int nipgpu_group_em_mstep(nipgpu_group* g, const double* counts) {
  if (!g) return gfail(NIPGPU_EINVAL, "bad arguments");
  return for_each_device(g, [&](int i) { return nipgpu_em_mstep(g->models[i], counts); });
}

}  // extern "C"
