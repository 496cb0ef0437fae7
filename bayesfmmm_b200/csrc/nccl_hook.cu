// nccl_hook.cu -- native all-reduce of the statistics buffer for function-sharded runs.
//
// The only exchange of a sharded sweep is the sum of a ~3 KB statistics buffer (twice per sweep: the whole
// buffer after the Z + statistics kernels, one double after the SSR pass).  The generic hook
// (bfmmm_sampler_set_allreduce) lets the caller plug any collective; through Python/torch.distributed it
// costs ~20 us per call, which is most of what a second GPU adds to a 0.36 ms sweep.  This file calls
// ncclAllReduce directly on the engine's stream.  NCCL is loaded with dlopen (the library torch already
// mapped, or an explicit path), so libbfmmm_b200.so has no link-time dependency on it; the unique id is
// produced on rank 0 and distributed by the caller (torch.distributed / MPI / a file).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <string>

#include "../../include/bfmmm_sampler.h"
#include "common.cuh"

namespace {
struct NcclId { char internal[128]; };
typedef void* ncclComm_t;
typedef int (*get_id_fn)(NcclId*);
typedef int (*init_rank_fn)(ncclComm_t*, int, NcclId, int);
typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*destroy_fn)(ncclComm_t);
typedef const char* (*errstr_fn)(int);
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;     // ncclFloat64, ncclSum (nccl.h)

struct Api {
  void* handle = nullptr;
  get_id_fn get_id = nullptr; init_rank_fn init_rank = nullptr; allreduce_fn allreduce = nullptr;
  destroy_fn destroy = nullptr; errstr_fn errstr = nullptr;
} g_api;

int load_api(const char* path) {
  if (g_api.handle) return 0;
  const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    g_api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.handle) break;
  }
  if (!g_api.handle) return bf::set_error("bfmmm_nccl: libnccl.so.2 not found (pass its path)");
  g_api.get_id = (get_id_fn)dlsym(g_api.handle, "ncclGetUniqueId");
  g_api.init_rank = (init_rank_fn)dlsym(g_api.handle, "ncclCommInitRank");
  g_api.allreduce = (allreduce_fn)dlsym(g_api.handle, "ncclAllReduce");
  g_api.destroy = (destroy_fn)dlsym(g_api.handle, "ncclCommDestroy");
  g_api.errstr = (errstr_fn)dlsym(g_api.handle, "ncclGetErrorString");
  if (!g_api.get_id || !g_api.init_rank || !g_api.allreduce || !g_api.destroy) {
    g_api.handle = nullptr;
    return bf::set_error("bfmmm_nccl: libnccl lacks the expected symbols");
  }
  return 0;
}
int nccl_fail(const char* what, int rc) {
  std::string m = std::string(what) + ": " + (g_api.errstr ? g_api.errstr(rc) : "NCCL error");
  return bf::set_error(m.c_str());
}
struct Ctx { ncclComm_t comm; };
int native_allreduce(void* ctx, double* buf, int64_t len, void* stream) {
  Ctx* c = (Ctx*)ctx;
  int rc = g_api.allreduce(buf, buf, (size_t)len, NCCL_DOUBLE, NCCL_SUM, c->comm, (cudaStream_t)stream);
  return rc == 0 ? 0 : nccl_fail("ncclAllReduce", rc);
}
}  // namespace

extern "C" {

// rank 0: produce the 128-byte unique id every rank must pass to bfmmm_sampler_enable_nccl
int bfmmm_nccl_unique_id(const char* libnccl_path, char* id_out /* 128 bytes */) {
  if (load_api(libnccl_path)) return 1;
  NcclId id;
  int rc = g_api.get_id(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  std::memcpy(id_out, id.internal, 128);
  return 0;
}

// collective over all ranks: creates the communicator (on the current CUDA device) and installs the
// native all-reduce as the sampler's hook; *comm_out must be released with bfmmm_nccl_destroy
int bfmmm_sampler_enable_nccl(bfmmm_sampler* s, const char* libnccl_path, const char* id /* 128 bytes */, int rank, int world,
                              void** comm_out) {
  if (!s || !id || !comm_out) return bf::set_error("bfmmm_sampler_enable_nccl: null argument");
  if (load_api(libnccl_path)) return 1;
  NcclId nid;
  std::memcpy(nid.internal, id, 128);
  Ctx* c = new Ctx();
  int rc = g_api.init_rank(&c->comm, world, nid, rank);
  if (rc) { delete c; return nccl_fail("ncclCommInitRank", rc); }
  *comm_out = c;
  return bfmmm_sampler_set_allreduce(s, native_allreduce, c);
}

void bfmmm_nccl_destroy(void* comm) {
  Ctx* c = (Ctx*)comm;
  if (!c) return;
  if (g_api.destroy) g_api.destroy(c->comm);
  delete c;
}

}  // extern "C"
