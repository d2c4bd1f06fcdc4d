// NCCL plumbing for the frame-sharded multi-GPU mode: ONE ncclAllGather per keyframe round replicates the new
// 512-d global descriptors into every rank's bank.  libnccl is dlopen'ed lazily so single-GPU use
// (and the CPU-side symbol tests) never need it; when the host process already loaded NCCL (torch), the same
// library instance is reused.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "engine.h"

namespace dv {

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
};

static Comm* g_api = nullptr;   // function table shared by all engines of the process

static int load_api() {
  if (g_api) return DV_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror()); return DV_ERR_COMM; }
  Comm* c = new Comm();
  c->lib = lib;
  c->GetUniqueId = reinterpret_cast<decltype(c->GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  c->CommInitRank = reinterpret_cast<decltype(c->CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  c->AllGather = reinterpret_cast<decltype(c->AllGather)>(dlsym(lib, "ncclAllGather"));
  c->CommDestroy = reinterpret_cast<decltype(c->CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  c->GetErrorString = reinterpret_cast<decltype(c->GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  c->CommGetAsyncError = reinterpret_cast<decltype(c->CommGetAsyncError)>(dlsym(lib, "ncclCommGetAsyncError"));
  c->CommAbort = reinterpret_cast<decltype(c->CommAbort)>(dlsym(lib, "ncclCommAbort"));
  if (!c->GetUniqueId || !c->CommInitRank || !c->AllGather || !c->CommDestroy) {
    delete c;
    set_error("libnccl is missing required symbols");
    return DV_ERR_COMM;
  }
  g_api = c;
  return DV_OK;
}

static int nccl_fail(const char* what, ncclResult_t r) {
  set_error(std::string(what) + ": " + (g_api && g_api->GetErrorString ? g_api->GetErrorString(r) : "nccl error"));
  return DV_ERR_COMM;
}

int comm_allgather(Engine* e, const float* send, float* recv, size_t count_per_rank) {
  if (!e->comm || !e->comm->comm) { set_error("world_size > 1 but dv_comm_init was not called"); return DV_ERR_COMM; }
  ncclResult_t r = g_api->AllGather(send, recv, count_per_rank, ncclFloat32, e->comm->comm, e->st);
  if (r != ncclSuccess) return nccl_fail("ncclAllGather", r);
  return DV_OK;
}

int comm_allgather_bytes(Engine* e, const void* send, void* recv, size_t bytes_per_rank) {
  if (!e->comm || !e->comm->comm) { set_error("world_size > 1 but dv_comm_init was not called"); return DV_ERR_COMM; }
  ncclResult_t r = g_api->AllGather(send, recv, bytes_per_rank, ncclInt8, e->comm->comm, e->st);
  if (r != ncclSuccess) return nccl_fail("ncclAllGather", r);
  return DV_OK;
}

// Failure detection: NCCL reports network / peer failures asynchronously; every collective of the path is
// followed (after its stream synchronisation) by this non-blocking query, so a dead peer surfaces as DV_ERR_COMM on the
// call that used the collective instead of a hang or silently stale bank rows.
int comm_poll(Engine* e) {
  if (!e->comm || !e->comm->comm || !g_api || !g_api->CommGetAsyncError) return DV_OK;
  ncclResult_t async = ncclSuccess;
  ncclResult_t r = g_api->CommGetAsyncError(e->comm->comm, &async);
  if (r != ncclSuccess) return nccl_fail("ncclCommGetAsyncError", r);
  if (async != ncclSuccess && async != ncclInProgress) {
    if (g_api->CommAbort) { g_api->CommAbort(e->comm->comm); e->comm->comm = nullptr; }
    return nccl_fail("NCCL asynchronous error (communicator aborted)", async);
  }
  return DV_OK;
}

void comm_free(Engine* e) {
  if (e->comm) {
    if (e->comm->comm && g_api) g_api->CommDestroy(e->comm->comm);
    delete e->comm;
    e->comm = nullptr;
  }
}

}  // namespace dv

using namespace dv;

extern "C" {

dv_status dv_comm_unique_id(void* id128) {
  if (!id128) { set_error("dv_comm_unique_id: null buffer"); return DV_ERR_INVALID; }
  DV_TRY(load_api());
  ncclUniqueId id;
  ncclResult_t r = g_api->GetUniqueId(&id);
  if (r != ncclSuccess) return (dv_status)nccl_fail("ncclGetUniqueId", r);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return DV_OK;
}

dv_status dv_comm_init(dv_engine* h, const void* id128) {
  if (!h || !id128) { set_error("dv_comm_init: null argument"); return DV_ERR_INVALID; }
  Engine* e = reinterpret_cast<Engine*>(h);
  if (e->cfg.world_size <= 1) return DV_OK;
  DV_TRY(load_api());
  DV_CUDA_OK(cudaSetDevice(e->cfg.device));
  comm_free(e);
  e->comm = new Comm();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_api->CommInitRank(&e->comm->comm, e->cfg.world_size, id, e->cfg.rank);
  if (r != ncclSuccess) { comm_free(e); return (dv_status)nccl_fail("ncclCommInitRank", r); }
  log_msg(3, "NCCL communicator up: rank %d of %d; exchanging CUDA-IPC handles of the feature stores", e->cfg.rank, e->cfg.world_size);
  return (dv_status)store_exchange_peers(e);     // CUDA-IPC views of every rank's feature store (one-sided P2P pulls)
}

}  // extern "C"
