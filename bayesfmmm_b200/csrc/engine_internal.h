// engine_internal.h -- entry points of engine.cu used by the sampler's device-resident sweep (host_sampler.cu).
// C++ linkage, inside libbfmmm_b200.so only: not part of the C ABI of include/bfmmm.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bfmmm.h"

namespace bf {
struct P2PPeers;
struct EngineDevInfo {
  double* stats;          // device statistics buffer [sum log Z (K) | accepts | ssr | ssr_after | W'W | C~'W ...]
  int64_t stats_len;
  double* glob;           // P4 x QS whitened global coefficients staged by the pass kernels
  int Pc, P4, QS, hbL;    // rows of the projected cache (rank of the basis Gram), padded rows, feature stride, band of L
  const double* L_host;   // P x Pc whitening factor (column-major, host)
  cudaStream_t stream;
  int device;
};
}  // namespace bf
int bfmmm_engine_devinfo(bfmmm_engine* e, bf::EngineDevInfo* out);
int bfmmm_update_z_async_p(bfmmm_engine* e, double a_Z_PM, double beta, const double* zpar_dev);
int bfmmm_update_chi_async_p(bfmmm_engine* e, double beta, const double* sigma_dev);
int bfmmm_set_sigma(bfmmm_engine* e, double sigma_sq);
void bfmmm_moments_invalidate(bfmmm_engine* e);   // the caller changed the staged globals behind the engine's back
int bfmmm_z_propose_async(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, uint64_t iteration, bool after_queued);
int bfmmm_slz_read_begin(bfmmm_engine* e);
int bfmmm_slz_read_wait(bfmmm_engine* e, double* out /* K + 1 */);
bool bfmmm_z_ahead_supported(bfmmm_engine* e);     // the Z step runs as proposal + accept kernels (common basis)
double bfmmm_z_propose_us(bfmmm_engine* e);        // last proposal kernel timed alone on the engine's stream (< 0: none yet)
// peer-memory exchange fused into the statistics pass's final reduction (csrc/p2p_hook.cu installs it)
int bfmmm_engine_set_exchange(bfmmm_engine* e, const bf::P2PPeers* peers, int rank, int world, int cap, unsigned long long* seq);
int bfmmm_stats_exchanged(bfmmm_engine* e);        // the statistics pass queued last summed the buffer over the shards itself
int64_t bfmmm_stats_len(bfmmm_engine* e);          // doubles in the statistics buffer
// SSR pass + sigma^2 draw (+ one-slot exchange) in one launch; *done = 0: not available, nothing launched
int bfmmm_ssr_sigma_async(bfmmm_engine* e, int need_exchange, double a_shape, double scale_ssr, double beta0, uint64_t key,
                          uint64_t iteration, uint32_t purpose, int* done);
