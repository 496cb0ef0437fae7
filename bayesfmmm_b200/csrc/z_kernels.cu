// z_kernels.cu -- instantiations + (K, M, covariates) dispatch of z_kernel (pass_kernels.cuh)
#include <cstdlib>

#include "pass_kernels.cuh"

namespace bf {
// functions per thread (V): see pass_kernels.cuh.  BF_TUNE_V builds both variants and lets the
// environment variable BFMMM_V_Z pick one at run time (tuning experiments only).
constexpr int KV = 1;   // measured on B200 (tools/kbench.py): Z 122 us with V = 1 vs 138 us with V = 2 (64-register builds)
#ifdef BF_TUNE_V
static int tune_v() { static int v = -1; if (v < 0) { const char* e = std::getenv("BFMMM_V_Z"); v = e ? std::atoi(e) : KV; } return v; }
#define BF_CASE_z(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    if (tune_v() == 2) return cov ? launch_pass<2>(z_kernel<KK, MM, true, 2, false>, a, s) : launch_pass<2>(z_kernel<KK, MM, false, 2, false>, a, s); \
    return cov ? launch_pass<1>(z_kernel<KK, MM, true, 1, false>, a, s) : launch_pass<1>(z_kernel<KK, MM, false, 1, false>, a, s);
#else
#define BF_CASE_z(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<KV>(z_kernel<KK, MM, true, KV, false>, a, s)      \
               : launch_pass<KV>(z_kernel<KK, MM, false, KV, false>, a, s);
#endif

int launch_z(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(z)
}

// the two halves on their own (see z_propose_kernel)
#ifndef BF_ZA_V
#define BF_ZA_V 2
#endif
#define BF_CASE_zaccept(KK, MM)                                                                  \
  case KK * 16 + MM:                                                                             \
    return cov ? launch_pass<BF_ZA_V>(z_kernel<KK, MM, true, BF_ZA_V, false, true>, a, s)        \
               : launch_pass<BF_ZA_V>(z_kernel<KK, MM, false, BF_ZA_V, false, true>, a, s);
int launch_z_accept(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(zaccept)
}
// Blocks of a few grid-stride iterations each, not one resident wave: the kernel usually runs on the low-priority side
// stream beside the sweep's own kernels, and short blocks hand their slots back to those as they finish.
template <int K>
static int launch_propose(const PassArgs& a, cudaStream_t s) {
  constexpr int ITER = 3;
  int grid = (a.ld + PF_THREADS * ITER - 1) / (PF_THREADS * ITER);
  if (grid < 1) grid = 1;
  z_propose_kernel<K><<<grid, PF_THREADS, 0, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}
int launch_z_propose(const PassArgs& a, int K, cudaStream_t s) {
  switch (K) {
    case 2: return launch_propose<2>(a, s);
    case 3: return launch_propose<3>(a, s);
    case 4: return launch_propose<4>(a, s);
    case 5: return launch_propose<5>(a, s);
    case 6: return launch_propose<6>(a, s);
    default: return -2;
  }
}
}  // namespace bf
