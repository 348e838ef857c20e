// Product-level multi-GPU support: one process per GPU, cells sharded by the host, and the path's ONE
// collective — the sum over all cells of the parameter sensitivities Mu (NP x NADJ doubles, 7.4 kB) — done inside the library
// over NCCL, so that a C or Fortran host can shard a batch without torch or MPI-side reductions.
//
//   rank 0:      fatode_comm_unique_id(id)            -> 128 bytes, handed to the other ranks by the host's own means
//   every rank:  fatode_comm_init(nranks, rank, id)   (after cudaSetDevice: the communicator binds to the current device)
//                ... fatode_*_integrate_adj_batch(..., mu_sum, ...): mu_sum comes back summed over ALL ranks' cells
//                fatode_comm_allreduce_sum(buf, n, mem, stream) for anything else the host wants summed (counters, ...)
//                fatode_comm_finalize()
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy the process already has — e.g. the one torch loaded — or the
// system's), so libfatode_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstring>
#include <string>
#include "../../include/fatode_b200.h"
#include "units_internal.h"

namespace fatode {
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(h, "ncclAllReduce"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl.so.2 lacks the expected entry points"; return false; }
    return true;
  }
};
NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_nranks = 1, g_rank = 0;
thread_local std::string g_comm_err;
}  // namespace

bool comm_active() { return g_comm != nullptr && g_nranks > 1; }

// in-place sum over the ranks of the communicator; d_buf is a DEVICE pointer; stream-ordered
int comm_allreduce_device(double* d_buf, size_t count, cudaStream_t st, std::string& err) {
  if (!g_comm) return 0;
  const ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, count, ncclFloat64, ncclSum, g_comm, st);
  if (r != 0) { err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return FATODE_ECUDA; }
  return 0;
}
}  // namespace fatode

using namespace fatode;

extern "C" const char* fatode_comm_last_error(void) { return g_comm_err.c_str(); }

extern "C" int fatode_comm_unique_id(void* id128) {
  if (!id128) return FATODE_EINVAL;
  if (!g_nccl.load(g_comm_err)) return FATODE_EUNSUPPORTED;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) { g_comm_err = "ncclGetUniqueId failed"; return FATODE_ECUDA; }
  std::memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int fatode_comm_init(int nranks, int rank, const void* id128) {
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return FATODE_EINVAL;
  if (g_comm) { g_comm_err = "communicator already initialised (fatode_comm_finalize first)"; return FATODE_EINVAL; }
  if (!g_nccl.load(g_comm_err)) return FATODE_EUNSUPPORTED;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  const ncclResult_t r = g_nccl.CommInitRank(&g_comm, nranks, id, rank);
  if (r != 0) {
    g_comm = nullptr;
    g_comm_err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return FATODE_ECUDA;
  }
  g_nranks = nranks;
  g_rank = rank;
  return 0;
}

extern "C" int fatode_comm_size(void) { return g_comm ? g_nranks : 1; }
extern "C" int fatode_comm_rank(void) { return g_comm ? g_rank : 0; }

extern "C" int fatode_comm_allreduce_sum(double* buf, int64_t count, int mem, void* stream) {
  if (!buf || count < 0) return FATODE_EINVAL;
  if (!g_comm || count == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mem == FATODE_MEM_DEVICE) return comm_allreduce_device(buf, (size_t)count, st, g_comm_err);
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * (size_t)count) != cudaSuccess) { g_comm_err = "cudaMalloc failed"; return FATODE_ECUDA; }
  int rc = 0;
  if (cudaMemcpyAsync(d, buf, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = FATODE_ECUDA;
  if (!rc) rc = comm_allreduce_device(d, (size_t)count, st, g_comm_err);
  if (!rc && cudaMemcpyAsync(buf, d, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = FATODE_ECUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = rc ? rc : FATODE_ECUDA;
  cudaFree(d);
  return rc;
}

extern "C" int fatode_comm_finalize(void) {
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_nranks = 1;
  g_rank = 0;
  return 0;
}
