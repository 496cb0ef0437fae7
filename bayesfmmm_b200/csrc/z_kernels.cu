// z_kernels.cu -- instantiations + (K, M, covariates) dispatch of z_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
// functions per thread: see pass_kernels.cuh (V = 1 doubles the resident warps of the
// latency-bound Z and chi kernels; the bandwidth-bound SSR pass keeps 16-byte accesses)
constexpr int Z_V = 1;
#define BF_CASE_z(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<Z_V>(z_kernel<KK, MM, true, Z_V>, a, s)      \
               : launch_pass<Z_V>(z_kernel<KK, MM, false, Z_V>, a, s);

int launch_z(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(z)
}
}  // namespace bf
