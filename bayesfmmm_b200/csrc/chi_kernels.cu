// chi_kernels.cu -- instantiations + (K, M, covariates) dispatch of chi_kernel (pass_kernels.cuh)
#include "pass_kernels.cuh"

namespace bf {
// functions per thread: see pass_kernels.cuh (V = 1 doubles the resident warps of the
// latency-bound Z and chi kernels; the bandwidth-bound SSR pass keeps 16-byte accesses)
constexpr int CHI_V = 1;
#define BF_CASE_chi(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<CHI_V>(chi_kernel<KK, MM, true, CHI_V>, a, s)      \
               : launch_pass<CHI_V>(chi_kernel<KK, MM, false, CHI_V>, a, s);

int launch_chi(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(chi)
}
}  // namespace bf
