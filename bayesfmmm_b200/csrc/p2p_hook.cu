// p2p_hook.cu -- one-shot all-reduce of the statistics buffer over NVLink peer memory.
//
// The buffer is ~3 KB and the ranks share one NVSwitch domain, so the exchange is latency, not
// bandwidth: ncclAllReduce costs ~14 us per call at two ranks (launch + protocol), twice per sweep.  Here
// every rank owns a "mailbox" in its own HBM, mapped into its peers with CUDA IPC; one small kernel per
// rank (a) stores its partial sums into slot [rank] of EVERY mailbox (remote stores over NVLink),
// (b) publishes a sequence number in each mailbox, (c) waits until all ranks' numbers have arrived in
// its own mailbox and (d) sums the slots in rank order -- the same order on every rank, so the result
// is bit-identical everywhere, which is what keeps the replicated host-side draws in lock step
// (SURVEY 8e).  Slots are double-buffered by the parity of the sequence number: a rank can start
// exchange s+1 while a peer still sums exchange s, and cannot reach s+2 before that peer has sent s+1.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/bfmmm_sampler.h"
#include "common.cuh"
#include "engine_internal.h"

bfmmm_engine* bfmmm_sampler_engine(bfmmm_sampler* s);      // host_sampler.cu

namespace {
using bf::P2P_MAX_RANKS;
using bf::P2P_HDR;
struct Mailbox {                 // layout of one rank's mailbox in its HBM
  unsigned long long flags[P2P_MAX_RANKS];     // sequence number last published by each rank
  // double data[2][world][cap] follows (256-byte aligned)
};
typedef bf::P2PPeers Peers;

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(double* __restrict__ buf, int len, Peers peers, int rank, int world,
                                                            unsigned long long seq, int cap) {
  const int par = (int)(seq & 1ull);
  // (a) my partial sums into slot [par][rank] of every mailbox
  for (int r = 0; r < world; r++) {
    double* dst = reinterpret_cast<double*>(peers.box[r] + P2P_HDR) + ((size_t)par * world + rank) * cap;
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  // (b) publish, (c) wait for everyone (bounded: a missing rank traps instead of hanging the device)
  if (threadIdx.x < world) {
    volatile unsigned long long* theirs = reinterpret_cast<Mailbox*>(peers.box[threadIdx.x])->flags + rank;
    *theirs = seq;
    volatile unsigned long long* mine = reinterpret_cast<Mailbox*>(peers.box[rank])->flags + threadIdx.x;
    unsigned long long spins = 0;
    while (*mine < seq) {
      if (++spins > (1ull << 31)) __trap();
    }
  }
  __threadfence_system();
  __syncthreads();
  // (d) sum in rank order
  const double* src = reinterpret_cast<const double*>(peers.box[rank] + P2P_HDR) + (size_t)par * world * cap;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    double t = 0;
    for (int r = 0; r < world; r++) t += __ldcg(src + (size_t)r * cap + i);
    buf[i] = t;
  }
}

struct Ctx {
  int rank = 0, world = 1, cap = 0, device = 0;
  unsigned long long seq = 0;
  unsigned char* local = nullptr;
  Peers peers;
  std::vector<void*> opened;
};

int p2p_allreduce(void* ctx, double* buf, int64_t len, void* stream) {
  Ctx* c = (Ctx*)ctx;
  if (len > c->cap) return bf::set_error("p2p all-reduce: buffer longer than the mailbox slots");
  c->seq++;
  p2p_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(buf, (int)len, c->peers, c->rank, c->world, c->seq, c->cap);
  bf::g_launch_count++;
  return cudaGetLastError() == cudaSuccess ? 0 : bf::set_error("p2p all-reduce: launch failed");
}
size_t box_bytes(int world, int cap) { return P2P_HDR + (size_t)2 * world * cap * sizeof(double); }
}  // namespace

extern "C" {

// Allocates this rank's mailbox (slots of `cap` doubles, cap >= the sampler's statistics length) on the
// current device and returns its CUDA IPC handle (64 bytes) for the other ranks.
int bfmmm_p2p_create(int rank, int world, int64_t cap, void** ctx_out, char* handle_out /* 64 bytes */) {
  if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || cap <= 0 || !ctx_out || !handle_out)
    return bf::set_error("bfmmm_p2p_create: bad argument (at most 16 ranks)");
  Ctx* c = new Ctx();
  c->rank = rank; c->world = world; c->cap = (int)((cap + 31) & ~31ll);
  cudaGetDevice(&c->device);
  const size_t bytes = box_bytes(world, c->cap);
  if (cudaMalloc(&c->local, bytes) != cudaSuccess || cudaMemset(c->local, 0, bytes) != cudaSuccess) {
    delete c;
    return bf::set_error("bfmmm_p2p_create: cudaMalloc failed");
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) {
    cudaFree(c->local); delete c;
    return bf::set_error("bfmmm_p2p_create: cudaIpcGetMemHandle failed");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::memcpy(handle_out, &h, 64);
  cudaDeviceSynchronize();
  *ctx_out = c;
  return 0;
}

// handles = world x 64 bytes (rank order, e.g. from an all-gather).  Maps the peers' mailboxes and installs
// the exchange as the sampler's all-reduce hook.  The caller must synchronise all ranks (a barrier) between
// this call and the first sweep, and before bfmmm_p2p_destroy.
int bfmmm_sampler_enable_p2p(bfmmm_sampler* s, void* ctx, const char* handles) {
  Ctx* c = (Ctx*)ctx;
  if (!s || !c || !handles) return bf::set_error("bfmmm_sampler_enable_p2p: null argument");
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) { c->peers.box[r] = c->local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return bf::set_error((std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)).c_str());
    c->opened.push_back(p);
    c->peers.box[r] = (unsigned char*)p;
  }
  if (bfmmm_sampler_set_allreduce(s, p2p_allreduce, c)) return 1;
  // The whole-buffer exchange behind the statistics pass runs inside that pass's final reduction and the SSR slot's
  // inside the SSR pass (stats_kernels.cu, moments_kernels.cu: same mailboxes, same sequence counter); the hook above
  // remains for the other one-slot exchanges (log-likelihood flush, models without the fused passes).
  if (bfmmm_engine* e = bfmmm_sampler_engine(s))
    if (bfmmm_engine_set_exchange(e, &c->peers, c->rank, c->world, c->cap, &c->seq)) return 1;
  return 0;
}

void bfmmm_p2p_destroy(void* ctx) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->opened) cudaIpcCloseMemHandle(p);
  cudaFree(c->local);
  delete c;
}

}  // extern "C"
