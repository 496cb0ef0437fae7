// z_kernels.cu -- instantiations + (K, M, covariates) dispatch of z_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
#define BF_CASE_z(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass(z_kernel<KK, MM, true>, a, s) : launch_pass(z_kernel<KK, MM, false>, a, s);

int launch_z(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(z)
}
}  // namespace bf
