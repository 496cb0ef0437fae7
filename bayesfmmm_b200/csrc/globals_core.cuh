// globals_core.cuh -- the prior updates of the driver loops (pi, alpha_3, tau, delta, A, gamma and the banded
// Gaussian block draw), written ONCE as __host__ __device__ templates: the host loop (host_sampler.cu) and the
// device-resident sweep (globals_kernels.cu) run the same statements on the same Philox streams, so the two chains
// coincide up to libm / libdevice rounding.
//
// Reference: UpdatePi.h:39-116, UpdateAlpha3.h:10-63, UpdateTau.h:18-68, UpdateDelta.h:17-66, UpdateGamma.h:17-38,
// UpdateA.h:17-135, Distributions.h:22-61; block draw: UpdateNu.h:64-69, UpdatePhi.h:72-82.
//
// Random numbers.  Every update draws from its own counter-based stream (seed, tick, purpose, element): one stream per
// independent unit of work (a gamma_{k,p,.} row, a delta_{k,.} chain, an A_{k,i}, a tau_k, a pi gamma, a block
// coefficient), so the units can run on different device threads and still reproduce the host's sequential loop.
// The template parameter R provides  `auto st = rng.open(purpose, element); st.gamma(a); st.normal(); st.uniform();`
// (StreamRng below; the host's tape of injected draws implements the same interface and ignores the element).
#pragma once
#include <math.h>

#include "../../include/bfmmm_sampler.h"
#include "common.cuh"

namespace bf {

// purposes of the global Philox streams
enum { HP_PI = 101, HP_ALPHA3, HP_PHI, HP_DELTA, HP_A, HP_GAMMA, HP_NU, HP_TAU, HP_SIGMA, HP_ETA, HP_XI,
       HP_TAU_ETA, HP_DELTA_XI, HP_A_XI, HP_GAMMA_XI, HP_TT };

struct StreamRng {
  uint64_t key, tick;
  __host__ __device__ RngStream open(uint32_t purpose, uint64_t element) const {
    return RngStream(key, 0xB200ull + (element << 16), tick, purpose);
  }
};

// the chain's global parameters (Armadillo layouts, see bfmmm_sampler.h) as plain pointers
struct GlobalsView {
  int K, P, M;
  int identity;                 // multivariate model: basis = identity, priors (1/tau) I
  int hbP;                      // half bandwidth of the penalty matrix
  double *nu, *Phi, *pi, *delta, *gamma, *A, *tau, *alpha3;
  const double* Pmat;           // P x P column-major (nullptr for the multivariate model)
  bfmmm_hyper h;
  double n_total;
  __host__ __device__ double& nu_(int k, int p) const { return nu[(size_t)p * K + k]; }
  __host__ __device__ double& Phi_(int k, int p, int m) const { return Phi[((size_t)m * P + p) * K + k]; }
  __host__ __device__ double& gamma_(int k, int p, int m) const { return gamma[((size_t)m * P + p) * K + k]; }
  __host__ __device__ double& delta_(int k, int m) const { return delta[(size_t)m * K + k]; }
  __host__ __device__ double& A_(int k, int i) const { return A[(size_t)i * K + k]; }
};

// ------------------------------------------------------------------ scalar helpers (host and device)
__host__ __device__ inline double g_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
__host__ __device__ inline double g_qnorm(double p) {          // Acklam's rational approximation + Halley refinement
  if (p <= 0) return -INFINITY;
  if (p >= 1) return INFINITY;
  const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                      1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                      6.680131188771972e+01, -1.328068155288572e+01};
  const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                      -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                      3.754408661907416e+00};
  double q, r, x;
  if (p < 0.02425) { q = sqrt(-2 * log(p));
    x = (((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((d[0]*q+d[1])*q+d[2])*q+d[3])*q+1);
  } else if (p <= 1 - 0.02425) { q = p - 0.5; r = q * q;
    x = (((((a[0]*r+a[1])*r+a[2])*r+a[3])*r+a[4])*r+a[5])*q / (((((b[0]*r+b[1])*r+b[2])*r+b[3])*r+b[4])*r+1);
  } else { q = sqrt(-2 * log(1 - p));
    x = -(((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((d[0]*q+d[1])*q+d[2])*q+d[3])*q+1);
  }
  for (int it = 0; it < 2; it++) {
    double e = g_pnorm(x) - p;
    double u = e * sqrt(2 * 3.14159265358979323846) * exp(x * x / 2);
    x = x - u / (1 + x * u / 2);
  }
  return x;
}
// truncated normal on [lo, inf): inverse-CDF draw from one uniform; log density
__host__ __device__ inline double g_rtruncnorm_lo(double mean, double sd, double lo, double u) {
  double pa = g_pnorm((lo - mean) / sd);
  return mean + sd * g_qnorm(pa + u * (1.0 - pa));
}
__host__ __device__ inline double g_dtruncnorm_lo_log(double x, double mean, double sd, double lo) {
  if (x < lo) return -INFINITY;
  const double LN_SQRT_2PI = 0.918938533204672741780329736406;
  double z = (x - mean) / sd;
  double scale = 1.0 - g_pnorm((lo - mean) / sd);
  return -(LN_SQRT_2PI + 0.5 * z * z + log(sd)) - log(scale);
}
__host__ __device__ inline double g_calc_lB(const double* al, int K) {      // Distributions.h:51-61
  double lB = 0, tot = 0;
  for (int k = 0; k < K; k++) { lB += lgamma(al[k]); tot += al[k]; }
  return lB - lgamma(tot);
}

// ------------------------------------------------------------------ pi (UpdatePi.h:84-116; lpdf_pi_PM :39-53)
// prop[k] = the K proposal gammas (already drawn: element k of HP_PI), u = the Metropolis uniform (element K)
__host__ __device__ inline void core_pi_accept(const GlobalsView& g, const double* slz, const double* gam, double u) {
  const int K = g.K;
  double prop[8], al[8], al2[8], sum = 0;
  for (int k = 0; k < K; k++) { al[k] = g.h.a_pi_PM * g.pi[k]; prop[k] = gam[k]; sum += gam[k]; }
  for (int k = 0; k < K; k++) prop[k] /= sum;
  const double a3 = *g.alpha3;
  double lnew = 0, lold = 0, ap[8], ao[8];
  for (int k = 0; k < K; k++) {
    lnew += (g.h.c[k] - 1) * log(prop[k]); lnew += (a3 * prop[k] - 1) * slz[k]; ap[k] = a3 * prop[k];
    lold += (g.h.c[k] - 1) * log(g.pi[k]); lold += (a3 * g.pi[k] - 1) * slz[k]; ao[k] = a3 * g.pi[k];
  }
  lnew -= g.n_total * g_calc_lB(ap, K);
  lold -= g.n_total * g_calc_lB(ao, K);
  double q_new = 0, q_old = 0;
  for (int k = 0; k < K; k++) { al2[k] = g.h.a_pi_PM * prop[k]; q_new += (al[k] - 1) * log(prop[k]); q_old += (al2[k] - 1) * log(g.pi[k]); }
  q_new -= g_calc_lB(al, K);
  q_old -= g_calc_lB(al2, K);
  const double acc = lnew - lold + q_old - q_new;
  if (log(u) < acc)
    for (int k = 0; k < K; k++) g.pi[k] = prop[k];
}
template <class R>
__host__ __device__ inline double core_pi_gamma(const GlobalsView& g, R& rng, int k) {
  double al = g.h.a_pi_PM * g.pi[k];
  double sh = al <= 0 ? 10.0 : al;                    // rdirichlet guard, Distributions.h:24-28
  auto st = rng.open(HP_PI, (uint64_t)k);
  return st.gamma(sh);
}
template <class R>
__host__ __device__ inline void core_update_pi(const GlobalsView& g, R& rng, const double* slz) {
  double gam[8];
  for (int k = 0; k < g.K; k++) gam[k] = core_pi_gamma(g, rng, k);
  auto st = rng.open(HP_PI, (uint64_t)g.K);
  core_pi_accept(g, slz, gam, st.uniform());
}

// ------------------------------------------------------------------ alpha_3 (UpdateAlpha3.h:36-63, lpdf_alpha3 :10-26)
template <class R>
__host__ __device__ inline void core_update_alpha3(const GlobalsView& g, R& rng, const double* slz) {
  const int K = g.K;
  auto st = rng.open(HP_ALPHA3, 0);
  const double sd = g.h.var_alpha3, cur = *g.alpha3;
  const double prop = g_rtruncnorm_lo(cur, sd, 0.0, st.uniform());
  double lv[2];
  for (int w = 0; w < 2; w++) {
    const double a3 = w == 0 ? cur : prop, a3_ph = w == 0 ? prop : cur;
    double l = (-g.h.b) * a3, ap[8];
    for (int k = 0; k < K; k++) { l += (a3 * g.pi[k] - 1) * slz[k]; ap[k] = a3 * g.pi[k]; }
    l -= g.n_total * g_calc_lB(ap, K);
    l += g_dtruncnorm_lo_log(a3_ph, a3_ph, sd, 0.0);      // as written in the reference (:23-24)
    lv[w] = l;
  }
  const double u = st.uniform();
  if (log(u) < lv[1] - lv[0]) *g.alpha3 = prop;
}

// ------------------------------------------------------------------ tau (UpdateTau.h:18-40; MV :47-68): integer division P / 2
template <class R>
__host__ __device__ inline void core_update_tau_k(const GlobalsView& g, R& rng, int k) {
  const int P = g.P;
  const double a = g.h.alpha_nu + (double)(P / 2);
  double quad = 0;
  for (int r = 0; r < P; r++) {
    double pr = 0;
    if (g.identity) pr = g.nu_(k, r);
    else {
      const int c0 = r - g.hbP > 0 ? r - g.hbP : 0, c1 = r + g.hbP < P - 1 ? r + g.hbP : P - 1;
      for (int c = c0; c <= c1; c++) pr += g.Pmat[(size_t)r * P + c] * g.nu_(k, c);      // banded, symmetric
    }
    quad += g.nu_(k, r) * pr;
  }
  const double b = g.h.beta_nu + 0.5 * quad;
  auto st = rng.open(HP_TAU, (uint64_t)k);
  const double gd = (1 / b) * st.gamma(a);
  g.tau[k] = g.identity ? 1 / gd : gd;
}

// ------------------------------------------------------------------ delta (UpdateDelta.h:17-66): the chain i = 0..M-1 of one k
// S[m] = sum_j gamma(k,j,m) Phi(k,j,m)^2, so that  p2 = 1 + 1/2 sum_m S[m] prod_{n <= m, n != i} delta(k,n)
template <class R>
__host__ __device__ inline void core_update_delta_k(const GlobalsView& g, R& rng, int k) {
  const int P = g.P, M = g.M;
  auto st = rng.open(HP_DELTA, (uint64_t)k);
  for (int i = 0; i < M; i++) {
    double p1, p2 = 1;
    if (i == 0) {
      p1 = g.A_(k, 0) + ((P * M) / 2.0);
      for (int j = 0; j < P; j++) {
        p2 += 0.5 * g.gamma_(k, j, 0) * (g.Phi_(k, j, 0) * g.Phi_(k, j, 0));
        for (int m = 1; m < M; m++) {
          double tt = 1;
          for (int nn = 1; nn <= m; nn++) tt *= g.delta_(k, nn);
          p2 += 0.5 * g.gamma_(k, j, m) * tt * (g.Phi_(k, j, m) * g.Phi_(k, j, m));
        }
      }
    } else {
      p1 = g.A_(k, 1) + ((P * (M - i)) / 2.0);
      for (int j = 0; j < P; j++)
        for (int m = i; m < M; m++) {
          double tt = 1;
          for (int nn = 0; nn <= m; nn++) if (nn != i) tt *= g.delta_(k, nn);
          p2 += 0.5 * g.gamma_(k, j, m) * tt * (g.Phi_(k, j, m) * g.Phi_(k, j, m));
        }
    }
    g.delta_(k, i) = (1 / p2) * st.gamma(p1);
  }
}

// ------------------------------------------------------------------ gamma (UpdateGamma.h:17-38): the row (i, l), j = 0..M-1
template <class R>
__host__ __device__ inline void core_update_gamma_row(const GlobalsView& g, R& rng, int i, int l) {
  const double nug = g.h.nu_1;
  auto st = rng.open(HP_GAMMA, (uint64_t)i * g.P + l);
  double ph = 1;
  for (int j = 0; j < g.M; j++) {
    ph *= g.delta_(i, j);
    const double scale = 2 / (nug + ph * (g.Phi_(i, l, j) * g.Phi_(i, l, j)));
    g.gamma_(i, l, j) = scale * st.gamma((nug + 1) / 2);
  }
}

// ------------------------------------------------------------------ A (UpdateA.h:58-135; lpdf_a1 :17-23, lpdf_a2 :33-44)
__host__ __device__ inline double g_lpdf_a1(double al, double be, double a, double delta) {
  return -log(tgamma(a)) + (a - 1) * log(delta) + (al - 1) * log(a) - (a * be);
}
__host__ __device__ inline double g_lpdf_a2(double al, double be, double a, const double* delta, int M, int stride) {
  const double x = M - 1;
  double l = -x * log(tgamma(a)) + (al - 1) * log(a) - (a * be);
  for (int i = 1; i < M; i++) l += (a - 1) * log(delta[(size_t)i * stride]);
  return l;
}
// one Metropolis step of a (first: a_1 with delta_row[0]; else a_2 with delta_row[1..M-1]); two uniforms from st
template <class S>
__host__ __device__ inline void core_mh_a(const bfmmm_hyper& h, S& st, double& a, bool first, const double* delta_row, int M, int stride) {
  const double sd = first ? h.var_epsilon1 / h.beta1l : h.var_epsilon2 / h.beta2l;
  const double cur = a;
  const double lp = first ? g_lpdf_a1(h.alpha1l, h.beta1l, cur, delta_row[0]) : g_lpdf_a2(h.alpha2l, h.beta2l, cur, delta_row, M, stride);
  const double prop = g_rtruncnorm_lo(cur, sd, 0.0, st.uniform());
  const double lpn = first ? g_lpdf_a1(h.alpha1l, h.beta1l, prop, delta_row[0]) : g_lpdf_a2(h.alpha2l, h.beta2l, prop, delta_row, M, stride);
  const double acc = (lpn + g_dtruncnorm_lo_log(cur, prop, sd, 0.0)) - lp - g_dtruncnorm_lo_log(prop, cur, sd, 0.0);
  const double u = st.uniform();
  if (log(u) < acc) a = prop;
}
template <class R>
__host__ __device__ inline void core_update_A_one(const GlobalsView& g, R& rng, int j, int i) {
  auto st = rng.open(HP_A, (uint64_t)j * 2 + i);
  core_mh_a(g.h, st, g.A_(j, i), i == 0, &g.delta_(j, 0), g.M, g.K);
}

// ------------------------------------------------------------------ banded reverse Cholesky and the block draw
// A = U U' with U upper triangular, both stored as bands: Ab[i * ldb + d] = A[i][i + d], Ub[i * ldb + d] = U[i][i + d],
// d = 0..hb (entries beyond the matrix are ignored).  See host_sampler.cu (chol_upper_rev) for why the reverse factor:
// chol_lower(A^-1) = U^-T, so the reference's draw  C b + chol_lower(C) z  is  U^-T (U^-1 b + z).
__host__ __device__ inline bool band_chol_upper_rev(int n, int hb, int ldb, const double* Ab, double* Ub) {
  for (int j = n - 1; j >= 0; j--) {
    // U[j][j]^2 = A[j][j] - sum_{k > j} U[j][k]^2 ;  U[i][j] = (A[i][j] - sum_{k > j} U[i][k] U[j][k]) / U[j][j]
    double s = Ab[(size_t)j * ldb];
    const int kmax = j + hb < n - 1 ? j + hb : n - 1;
    for (int k = j + 1; k <= kmax; k++) { const double u = Ub[(size_t)j * ldb + (k - j)]; s -= u * u; }
    if (!(s > 0)) return false;
    const double d = sqrt(s);
    Ub[(size_t)j * ldb] = d;
    const int imin = j - hb > 0 ? j - hb : 0;
    for (int i = j - 1; i >= imin; i--) {
      double t = Ab[(size_t)i * ldb + (j - i)];
      const int km = i + hb < n - 1 ? i + hb : n - 1;
      for (int k = j + 1; k <= km; k++) t -= Ub[(size_t)i * ldb + (k - i)] * Ub[(size_t)j * ldb + (k - j)];
      Ub[(size_t)i * ldb + (j - i)] = t / d;
    }
  }
  return true;
}
// the same factorisation that also returns rd[j] = 1 / U[j][j] and multiplies by it instead of dividing (one
// reciprocal square root per column: rsqrt on the device)
__host__ __device__ inline bool band_chol_upper_rev_rd(int n, int hb, int ldb, const double* Ab, double* Ub, double* rd) {
  for (int j = n - 1; j >= 0; j--) {
    double s = Ab[(size_t)j * ldb];
    const int kmax = j + hb < n - 1 ? j + hb : n - 1;
    for (int k = j + 1; k <= kmax; k++) { const double u = Ub[(size_t)j * ldb + (k - j)]; s -= u * u; }
    if (!(s > 0)) return false;
#ifdef __CUDA_ARCH__
    const double inv = rsqrt(s);
#else
    const double inv = 1.0 / sqrt(s);
#endif
    Ub[(size_t)j * ldb] = s * inv;
    rd[j] = inv;
    const int imin = j - hb > 0 ? j - hb : 0;
    for (int i = j - 1; i >= imin; i--) {
      double t = Ab[(size_t)i * ldb + (j - i)];
      const int km = i + hb < n - 1 ? i + hb : n - 1;
      for (int k = j + 1; k <= km; k++) t -= Ub[(size_t)i * ldb + (k - i)] * Ub[(size_t)j * ldb + (k - j)];
      Ub[(size_t)i * ldb + (j - i)] = t * inv;
    }
  }
  return true;
}
// x = U^-T (U^-1 b + z); w is scratch of length n
__host__ __device__ inline void band_draw(int n, int hb, int ldb, const double* Ub, const double* b, const double* z, double* x, double* w) {
  for (int i = n - 1; i >= 0; i--) {                 // U w = b
    double s = b[i];
    const int ki = i + hb < n - 1 ? i + hb : n - 1;
    for (int k = i + 1; k <= ki; k++) s -= Ub[(size_t)i * ldb + (k - i)] * w[k];
    w[i] = s / Ub[(size_t)i * ldb];
  }
  for (int i = 0; i < n; i++) w[i] += z[i];
  for (int i = 0; i < n; i++) {                      // U' x = w  (row-oriented: x_i = (w_i - sum_{k < i} U[k][i] x_k) / U[i][i])
    double s = w[i];
    const int k0 = i - hb > 0 ? i - hb : 0;
    for (int k = k0; k < i; k++) s -= Ub[(size_t)k * ldb + (i - k)] * x[k];
    x[i] = s / Ub[(size_t)i * ldb];
  }
}

// the same with the reciprocals rd[i] = 1 / U[i][i] precomputed (the sequential solves multiply instead of dividing)
__host__ __device__ inline void band_draw_rd(int n, int hb, int ldb, const double* Ub, const double* rd, const double* b,
                                             const double* z, double* x, double* w) {
  for (int i = n - 1; i >= 0; i--) {                 // U w = b
    double s = b[i];
    const int ki = i + hb < n - 1 ? i + hb : n - 1;
    for (int k = i + 1; k <= ki; k++) s -= Ub[(size_t)i * ldb + (k - i)] * w[k];
    w[i] = s * rd[i];
  }
  for (int i = 0; i < n; i++) {                      // U' x = w + z
    double s = w[i] + z[i];
    const int k0 = i - hb > 0 ? i - hb : 0;
    for (int k = k0; k < i; k++) s -= Ub[(size_t)k * ldb + (i - k)] * x[k];
    x[i] = s * rd[i];
  }
}

}  // namespace bf
