// chi_kernels_ragged.cu -- instantiations + (K, M, covariates) dispatch of chi_kernel (pass_kernels.cuh), ragged-grid variant
#include "pass_kernels.cuh"

namespace bf {
// functions per thread: see pass_kernels.cuh (V = 1 doubles the resident warps of the
// latency-bound Z and chi kernels; the bandwidth-bound SSR pass keeps 16-byte accesses)
constexpr int KV = 1;
constexpr bool KRG = true;
#define BF_CASE_chi(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<KV>(chi_kernel<KK, MM, true, KV, KRG>, a, s)      \
               : launch_pass<KV>(chi_kernel<KK, MM, false, KV, KRG>, a, s);

int launch_chi_ragged(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(chi)
}
}  // namespace bf
