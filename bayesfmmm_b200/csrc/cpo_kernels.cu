// cpo_kernels.cu -- instantiations + dispatch of the marginal log-likelihood / CPO mode of chi_kernel
// (pass_kernels.cuh, CPO = true): common and ragged grids.  Post-processing path (SURVEY 8f, f4:
// calcLikelihoodCPO, CalculateLikelihood.h:344-385), not part of a sampler sweep.
#include "pass_kernels.cuh"

namespace bf {
#define BF_CASE_cpo(KK, MM)                                                                          \
  case KK * 16 + MM:                                                                               \
    if (ragged) return cov ? launch_pass<1>(chi_kernel<KK, MM, true, 1, true, true>, a, s)          \
                           : launch_pass<1>(chi_kernel<KK, MM, false, 1, true, true>, a, s);        \
    return cov ? launch_pass<1>(chi_kernel<KK, MM, true, 1, false, true>, a, s)                     \
               : launch_pass<1>(chi_kernel<KK, MM, false, 1, false, true>, a, s);

int launch_mloglik(const PassArgs& a, int K, int M, bool ragged, cudaStream_t s) {
  BF_DISPATCH(cpo)
}
}  // namespace bf
