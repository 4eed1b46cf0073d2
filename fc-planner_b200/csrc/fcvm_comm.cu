// fcvm_comm.cu — native NCCL collectives for the sharded manager (see fcvm_comm.h).
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <string>

#include "fcvm_comm.h"
#include "fcvm_hostutil.h"

namespace fcvm {

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err, version;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    // a process that already holds a libnccl.so.2 (PyTorch bundles one) gets that very copy back: same soname
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) { a.err = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
    auto sym = [&](const char* n) { void* p = dlsym(a.lib, n); if (!p) a.err = std::string("libnccl lacks ") + n; return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    a.ok = a.err.empty();
    int v = 0;
    if (a.ok && a.GetVersion(&v) == ncclSuccess) a.version = std::to_string(v / 10000) + "." + std::to_string(v / 100 % 100) + "." + std::to_string(v % 100);
  });
  return a;
}

int nccl_fail(const char* what, ncclResult_t r) {
  return fail(FCVM_ECUDA, std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(r) : "NCCL error"));
}
}  // namespace

struct NcclComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

const char* nccl_version_string() { return api().ok ? api().version.c_str() : ""; }

int nccl_unique_id(uint8_t* id128) {
  NcclApi& a = api();
  if (!a.ok) return fail(FCVM_ENODEVICE, a.err);
  ncclUniqueId id;
  const ncclResult_t r = a.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, sizeof(id));
  return FCVM_OK;
}

int nccl_comm_create(const uint8_t* id128, int rank, int world, NcclComm** out) {
  NcclApi& a = api();
  if (!a.ok) return fail(FCVM_ENODEVICE, a.err);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NcclComm* c = new NcclComm();
  c->rank = rank; c->world = world;
  const ncclResult_t r = a.CommInitRank(&c->comm, world, id, rank);   // blocks until every rank has joined
  if (r != ncclSuccess) { delete c; return nccl_fail("ncclCommInitRank", r); }
  *out = c;
  return FCVM_OK;
}

void nccl_comm_destroy(NcclComm* c) {
  if (!c) return;
  if (c->comm && api().ok) api().CommDestroy(c->comm);
  delete c;
}

int nccl_comm_call(void* user, int32_t op, void* buf_dev, int64_t n, void* stream) {
  NcclComm* c = static_cast<NcclComm*>(user);
  NcclApi& a = api();
  if (!c || !a.ok) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ncclResult_t r;
  if (op == FCVM_COMM_ALLGATHER) {
    const char* mine = static_cast<const char*>(buf_dev) + (size_t)c->rank * (size_t)n;
    r = a.AllGather(mine, buf_dev, (size_t)n, ncclInt8, c->comm, s);     // in place: slice `rank` of the receive buffer is the send buffer
  } else if (op == FCVM_COMM_ALLREDUCE_MAX) {
    r = a.AllReduce(buf_dev, buf_dev, (size_t)n, ncclInt64, ncclMax, c->comm, s);
  } else {
    return -1;
  }
  if (r != ncclSuccess) { nccl_fail(op == FCVM_COMM_ALLGATHER ? "ncclAllGather" : "ncclAllReduce", r); return -1; }
  return 0;
}

}  // namespace fcvm
