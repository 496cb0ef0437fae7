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
}  // namespace bf
