// ssr_kernels.cu -- instantiations + (K, M, covariates) dispatch of ssr_kernel (pass_kernels.cuh)
#include <cstdlib>

#include "pass_kernels.cuh"

namespace bf {
// functions per thread (V): see pass_kernels.cuh.  BF_TUNE_V builds both variants and lets the
// environment variable BFMMM_V_SSR pick one at run time (tuning experiments only).
constexpr int KV = 2;
#ifdef BF_TUNE_V
static int tune_v() { static int v = -1; if (v < 0) { const char* e = std::getenv("BFMMM_V_SSR"); v = e ? std::atoi(e) : KV; } return v; }
#define BF_CASE_ssr(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    if (tune_v() == 2) return cov ? launch_pass<2>(ssr_kernel<KK, MM, true, 2, false>, a, s) : launch_pass<2>(ssr_kernel<KK, MM, false, 2, false>, a, s); \
    return cov ? launch_pass<1>(ssr_kernel<KK, MM, true, 1, false>, a, s) : launch_pass<1>(ssr_kernel<KK, MM, false, 1, false>, a, s);
#else
#define BF_CASE_ssr(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<KV>(ssr_kernel<KK, MM, true, KV, false>, a, s)      \
               : launch_pass<KV>(ssr_kernel<KK, MM, false, KV, false>, a, s);
#endif

int launch_ssr(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(ssr)
}
}  // namespace bf
