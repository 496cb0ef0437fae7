// chi_kernels.cu -- instantiations + (K, M, covariates) dispatch of chi_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
#define BF_CASE_chi(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass(chi_kernel<KK, MM, true>, a, s) : launch_pass(chi_kernel<KK, MM, false>, a, s);

int launch_chi(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(chi)
}
}  // namespace bf
