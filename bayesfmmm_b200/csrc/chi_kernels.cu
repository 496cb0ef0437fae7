// chi_kernels.cu -- instantiations + (K, M, covariates) dispatch of chi_kernel (pass_kernels.cuh)
#include <cstdlib>

#include "pass_kernels.cuh"

namespace bf {
// functions per thread (V): see pass_kernels.cuh.  BF_TUNE_V builds both variants and lets the
// environment variable BFMMM_V_CHI pick one at run time (tuning experiments only).
constexpr int KV = 2;   // measured on B200 (tools/kbench.py): chi 66 us with V = 2 vs 72 us with V = 1
#ifdef BF_TUNE_V
static int tune_v() { static int v = -1; if (v < 0) { const char* e = std::getenv("BFMMM_V_CHI"); v = e ? std::atoi(e) : KV; } return v; }
#define BF_CASE_chi(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    if (tune_v() == 2) return cov ? launch_pass<2>(chi_kernel<KK, MM, true, 2, false>, a, s) : launch_pass<2>(chi_kernel<KK, MM, false, 2, false>, a, s); \
    return cov ? launch_pass<1>(chi_kernel<KK, MM, true, 1, false>, a, s) : launch_pass<1>(chi_kernel<KK, MM, false, 1, false>, a, s);
#else
#define BF_CASE_chi(KK, MM)                                                          \
  case KK * 16 + MM:                                                               \
    return cov ? launch_pass<KV>(chi_kernel<KK, MM, true, KV, false>, a, s)      \
               : launch_pass<KV>(chi_kernel<KK, MM, false, KV, false>, a, s);
#endif

int launch_chi(const PassArgs& a, int K, int M, cudaStream_t s) {
  BF_DISPATCH(chi)
}
}  // namespace bf
