// ssr_kernels.cu -- instantiations + (K, M, covariates) dispatch of ssr_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
#define BF_CASE_ssr(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass(ssr_kernel<KK, MM, true>, a, s) : launch_pass(ssr_kernel<KK, MM, false>, a, s);

int launch_ssr(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(ssr)
}
}  // namespace bf
