// nrv_comm_*: the gradient all-reduce of data-parallel training, issued from the library.
//
// The reference delegates data parallelism to an external trainer that wraps the model in torch DDP
// (examples/evaluation.py:137-138, per-rank batch = batch_size // world_size, examples/CIFAR100.py:22).  Here the fused
// backward reports finished stages and parallel.py launches one in-place SUM all-reduce per bucket of the flat fp32
// gradient buffer on a side stream; this file owns the NCCL communicator those calls go through:
//   * ncclCommInitRankConfig with maxCTAs: the backward kernels are persistent and fill every SM, so an all-reduce that
//     asks for many CTAs only queues behind them -- a few CTAs (NVLink 5 / NVSwitch: the copy engines of the switch do
//     the reduction when NVLS is available) keep the overlap cheap;
//   * the flat gradient buffer is registered once (ncclCommRegister), which lets NCCL use it in place for its
//     NVLS / zero-copy paths instead of staging through its own buffers.
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries -- torch loads it), so libnrvit.so
// has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include "nrvit_internal.h"

#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <string.h>

namespace nrv {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommRegister)(const ncclComm_t, void*, size_t, void**) = nullptr;
  ncclResult_t (*CommDeregister)(const ncclComm_t, void*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_bind() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle != nullptr) return NRV_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy torch.distributed already loaded
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    set_error("nrv_comm: libnccl.so.2 not found (%s); initialise torch.distributed with the nccl backend first", dlerror());
    return NRV_ENOTIMPL;
  }
#define NRV_SYM(field, name)                                                  \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                  \
  if (g_nccl.field == nullptr) {                                              \
    set_error("nrv_comm: symbol %s missing in libnccl", name);                \
    return NRV_ENOTIMPL;                                                      \
  }
  NRV_SYM(GetUniqueId, "ncclGetUniqueId");
  NRV_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
  NRV_SYM(AllReduce, "ncclAllReduce");
  NRV_SYM(CommRegister, "ncclCommRegister");
  NRV_SYM(CommDeregister, "ncclCommDeregister");
  NRV_SYM(CommDestroy, "ncclCommDestroy");
  NRV_SYM(GetVersion, "ncclGetVersion");
  NRV_SYM(GetErrorString, "ncclGetErrorString");
#undef NRV_SYM
  g_nccl.handle = h;
  return NRV_OK;
}

#define NRV_NCCL(call)                                                                         \
  do {                                                                                         \
    ncclResult_t _r = (call);                                                                  \
    if (_r != ncclSuccess) {                                                                   \
      set_error("nrv_comm: %s failed: %s", #call, g_nccl.GetErrorString(_r));                   \
      return NRV_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

}  // namespace nrv

using namespace nrv;

struct nrv_comm {
  ncclComm_t comm;
  int nranks, rank;
};

extern "C" {

int nrv_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

int nrv_comm_get_unique_id(void* id_out, int bytes) {
  int rc = nccl_bind();
  if (rc) return rc;
  NRV_REQUIRE(id_out != nullptr && bytes >= (int)sizeof(ncclUniqueId), "nrv_comm_get_unique_id: buffer of nrv_comm_unique_id_bytes() needed");
  ncclUniqueId id;
  NRV_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return NRV_OK;
}

int nrv_comm_init(const void* id, int bytes, int nranks, int rank, int max_ctas, nrv_comm** out) {
  int rc = require_init();
  if (rc) return rc;
  rc = nccl_bind();
  if (rc) return rc;
  NRV_REQUIRE(id != nullptr && bytes >= (int)sizeof(ncclUniqueId) && out != nullptr, "nrv_comm_init: null pointer / short id");
  NRV_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "nrv_comm_init: bad rank %d of %d", rank, nranks);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  cfg.blocking = 1;
  if (max_ctas > 0) { cfg.minCTAs = 1; cfg.maxCTAs = max_ctas; }
  nrv_comm* c = new nrv_comm();
  c->nranks = nranks; c->rank = rank;
  ncclResult_t r = g_nccl.CommInitRankConfig(&c->comm, nranks, uid, rank, &cfg);
  if (r != ncclSuccess) {
    set_error("nrv_comm_init: ncclCommInitRankConfig failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return NRV_ECUDA;
  }
  *out = c;
  return NRV_OK;
}

int nrv_comm_register(nrv_comm* c, void* buf, size_t bytes, void** handle) {
  NRV_REQUIRE(c && buf && handle, "nrv_comm_register: null pointer");
  NRV_NCCL(g_nccl.CommRegister(c->comm, buf, bytes, handle));
  return NRV_OK;
}

int nrv_comm_deregister(nrv_comm* c, void* handle) {
  NRV_REQUIRE(c && handle, "nrv_comm_deregister: null pointer");
  NRV_NCCL(g_nccl.CommDeregister(c->comm, handle));
  return NRV_OK;
}

int nrv_comm_allreduce_bucket(nrv_comm* c, void* buf, long long count, int dtype, void* stream) {
  NRV_REQUIRE(c && buf && count >= 0, "nrv_comm_allreduce_bucket: null pointer");
  NRV_REQUIRE(dtype == NRV_F32 || dtype == NRV_BF16, "nrv_comm_allreduce_bucket: dtype must be NRV_F32 or NRV_BF16");
  if (count == 0) return NRV_OK;
  NRV_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, dtype == NRV_F32 ? ncclFloat32 : ncclBfloat16, ncclSum, c->comm,
                            (cudaStream_t)stream));
  return NRV_OK;
}

int nrv_comm_nccl_version(void) {
  if (nccl_bind()) return 0;
  int v = 0;
  g_nccl.GetVersion(&v);
  return v;
}

int nrv_comm_destroy(nrv_comm* c) {
  if (c == nullptr) return NRV_OK;
  if (g_nccl.handle != nullptr) g_nccl.CommDestroy(c->comm);
  delete c;
  return NRV_OK;
}

}  // extern "C"
