// Stand-in for RcppDist's <truncnorm.h> (r_truncnorm / d_truncnorm), used by
// UpdateAlpha3.h:23,45 and UpdateA.h:79-87.  TEST INFRASTRUCTURE ONLY.
// r_truncnorm draws by inverse CDF from ONE uniform (popped from the shim tape), which is
// distributionally the same as RcppDist's rejection sampler but not stream-identical
// (RcppDist is not vendored in the reference: "parity unpinned" for this draw).
#ifndef BFMMM_SHIM_TRUNCNORM_H
#define BFMMM_SHIM_TRUNCNORM_H
#include <RcppArmadillo.h>

namespace shim {
inline double pnorm_std(double x) { return 0.5 * std::erfc(-x / std::sqrt(2.0)); }
inline double qnorm_std(double p) {
  if (p <= 0) return -std::numeric_limits<double>::infinity();
  if (p >= 1) return std::numeric_limits<double>::infinity();
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  double q, r, x;
  if (p < 0.02425) { q = std::sqrt(-2 * std::log(p));
    x = (((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((d[0]*q+d[1])*q+d[2])*q+d[3])*q+1);
  } else if (p <= 1 - 0.02425) { q = p - 0.5; r = q * q;
    x = (((((a[0]*r+a[1])*r+a[2])*r+a[3])*r+a[4])*r+a[5])*q / (((((b[0]*r+b[1])*r+b[2])*r+b[3])*r+b[4])*r+1);
  } else { q = std::sqrt(-2 * std::log(1 - p));
    x = -(((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((d[0]*q+d[1])*q+d[2])*q+d[3])*q+1);
  }
  for (int it = 0; it < 2; it++) {   // Halley refinement to full double precision
    double e = pnorm_std(x) - p;
    double u = e * std::sqrt(2 * 3.14159265358979323846) * std::exp(x * x / 2);
    x = x - u / (1 + x * u / 2);
  }
  return x;
}
}  // namespace shim

inline double r_truncnorm(double mean, double sd, double a, double b) {
  double pa = shim::pnorm_std((a - mean) / sd), pb = shim::pnorm_std((b - mean) / sd);
  double u = shim::unif_rand();
  return mean + sd * shim::qnorm_std(pa + u * (pb - pa));
}
inline double d_truncnorm(double x, double mean, double sd, double a, double b, int lg) {
  if (x < a || x > b) return lg ? -std::numeric_limits<double>::infinity() : 0.0;
  double scale = shim::pnorm_std((b - mean) / sd) - shim::pnorm_std((a - mean) / sd);
  double l = R::dnorm(x, mean, sd, 1) - std::log(scale);
  return lg ? l : std::exp(l);
}
#endif
