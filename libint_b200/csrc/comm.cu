// NCCL plumbing behind the C ABI: the sum of the ranks' partial Fock matrices
// (the reference sums thread-private G's on the host, hartree-fock++.cc:1753-1755; here one
// ncclAllReduce(sum, f64) over NVLink per build).  A C++ host finishes a sharded build with
//   lb200_fock_build(f, D, ..., rank, nranks, G_dev, 1, ...);  lb200_fock_allreduce(comm, G_dev, nbf*nbf);
// without torch.  NCCL is bound at run time (dlopen): inside a torch process the already loaded
// libnccl.so.2 is reused, so there is never a second NCCL in the address space; the library itself
// does not link against NCCL and single-GPU users do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "internal.h"

using namespace lb200;

struct lb200_comm {
  lb200_context* ctx = nullptr;
  ncclComm_t comm = nullptr;
  bool owned = false;
  int rank = 0, nranks = 1;
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {std::getenv("LB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);   // prefer the copy the process already has
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
    api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.CommCount &&
             api.CommUserRank && api.GetErrorString;
  });
  return api;
}

int check_nccl(const lb200_context* ctx, ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return LB200_OK;
  return set_error(ctx, LB200_ERR_CUDA, std::string(what) + ": " + nccl().GetErrorString(r));
}

}  // namespace

extern "C" {

int lb200_comm_unique_id(char* id, int cap) {
  if (!id || cap < (int)sizeof(ncclUniqueId)) return LB200_ERR_INVALID;
  if (!nccl().ok) return LB200_ERR_CUDA;
  ncclUniqueId u;
  if (nccl().GetUniqueId(&u) != ncclSuccess) return LB200_ERR_CUDA;
  std::memcpy(id, &u, sizeof(u));
  return (int)sizeof(u);
}

int lb200_comm_create(lb200_context* ctx, int nranks, int rank, const char* id, lb200_comm** out) {
  if (!ctx || !id || !out || nranks < 1 || rank < 0 || rank >= nranks) return LB200_ERR_INVALID;
  if (!nccl().ok) return set_error(ctx, LB200_ERR_CUDA, "NCCL library (libnccl.so.2) not found");
  cudaSetDevice(ctx->device);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  auto* c = new lb200_comm;
  c->ctx = ctx; c->owned = true; c->rank = rank; c->nranks = nranks;
  int rc = check_nccl(ctx, nccl().CommInitRank(&c->comm, nranks, u, rank), "ncclCommInitRank");
  if (rc) { delete c; return rc; }
  *out = c;
  return LB200_OK;
}

int lb200_comm_from_nccl(lb200_context* ctx, void* nccl_comm, lb200_comm** out) {
  if (!ctx || !nccl_comm || !out) return LB200_ERR_INVALID;
  if (!nccl().ok) return set_error(ctx, LB200_ERR_CUDA, "NCCL library (libnccl.so.2) not found");
  auto* c = new lb200_comm;
  c->ctx = ctx; c->comm = static_cast<ncclComm_t>(nccl_comm); c->owned = false;
  int rc = check_nccl(ctx, nccl().CommCount(c->comm, &c->nranks), "ncclCommCount");
  if (!rc) rc = check_nccl(ctx, nccl().CommUserRank(c->comm, &c->rank), "ncclCommUserRank");
  if (rc) { delete c; return rc; }
  *out = c;
  return LB200_OK;
}

int lb200_comm_destroy(lb200_comm* c) {
  if (!c) return LB200_OK;
  if (c->owned && c->comm) {
    cudaSetDevice(c->ctx->device);
    nccl().CommDestroy(c->comm);
  }
  delete c;
  return LB200_OK;
}

int lb200_comm_rank(const lb200_comm* c) { return c ? c->rank : LB200_ERR_INVALID; }
int lb200_comm_size(const lb200_comm* c) { return c ? c->nranks : LB200_ERR_INVALID; }

int lb200_fock_allreduce(lb200_comm* c, double* G_device, long long count) {
  if (!c || !G_device || count < 0) return LB200_ERR_INVALID;
  cudaSetDevice(c->ctx->device);
  return check_nccl(c->ctx, nccl().AllReduce(G_device, G_device, (size_t)count, ncclDouble, ncclSum, c->comm,
                                             c->ctx->stream), "ncclAllReduce");
}

}  // extern "C"
