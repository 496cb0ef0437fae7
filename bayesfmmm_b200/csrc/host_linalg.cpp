// host_linalg.cpp -- the one host loop that matters at large P: the band-limited reverse Cholesky of the
// block draws (host_sampler.cu).  Plain C++ compiled by the host compiler so that an AVX2+FMA clone of the
// same source can be selected at run time (the inner loops are contiguous dot products of up to `hb` terms);
// machines without AVX2 take the baseline clone.  Rounding differs between the clones in the last bits only.
#include <algorithm>
#include <cmath>
#include <cstddef>

namespace bf_host {

#define BF_CHOL_BODY                                                              \
  for (int j = n - 1; j >= 0; j--) {                                              \
    double* uj = U + (size_t)j * n;                                               \
    const int kj = std::min(n - 1, j + hb);                                       \
    double s = A[(size_t)j * n + j];                                              \
    for (int k = j + 1; k <= kj; k++) s -= uj[k] * uj[k];                         \
    if (!(s > 0)) return false;                                                   \
    const double d = std::sqrt(s);                                                \
    uj[j] = d;                                                                    \
    for (int i = std::max(0, j - hb); i < j; i++) {                               \
      double* ui = U + (size_t)i * n;                                             \
      double t = A[(size_t)j * n + i];                                            \
      const int ki = std::min(n - 1, i + hb);                                     \
      for (int k = j + 1; k <= ki; k++) t -= ui[k] * uj[k];                       \
      ui[j] = t / d;                                                              \
    }                                                                             \
  }                                                                               \
  return true;

static bool chol_upper_rev_base(int n, const double* A, double* U, int hb) { BF_CHOL_BODY }

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2,fma"), optimize("O3,associative-math,no-signed-zeros,no-trapping-math")))
static bool chol_upper_rev_avx2(int n, const double* A, double* U, int hb) { BF_CHOL_BODY }
#endif

// A = U U' with U upper triangular (row-major rows), A banded with half bandwidth hb; false if not positive definite
bool chol_upper_rev(int n, const double* A, double* U, int hb) {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (avx2 && n >= 64) return chol_upper_rev_avx2(n, A, U, hb);
#endif
  return chol_upper_rev_base(n, A, U, hb);
}

}  // namespace bf_host
