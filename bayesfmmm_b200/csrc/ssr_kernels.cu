// ssr_kernels.cu -- instantiations + (K, M, covariates) dispatch of ssr_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
// functions per thread: see pass_kernels.cuh (V = 1 doubles the resident warps of the
// latency-bound Z and chi kernels; the bandwidth-bound SSR pass keeps 16-byte accesses)
constexpr int SSR_V = 2;
#define BF_CASE_ssr(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<SSR_V>(ssr_kernel<KK, MM, true, SSR_V>, a, s)      \
               : launch_pass<SSR_V>(ssr_kernel<KK, MM, false, SSR_V>, a, s);

int launch_ssr(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(ssr)
}
}  // namespace bf
