// fastmath.cuh -- short straight-line FP64 transcendentals for the per-function kernels.
//
// The Z step needs, per function, ~7 logarithms, 2 square roots, 2 sine/cosine pairs, ~10 reciprocals
// and 2(K+1) log-Gammas (reference UpdateMixedMembership.h:20-50,102-113; Distributions.h:22-61).
// libdevice's versions cost 60-200 instructions each, most of them not FP64 math (special-case
// branches, 64-bit immediates materialised two UMOVs at a time, call sequences): the first profile of
// the Z kernel showed 3500 instructions per function of which 1130 were FP64 (profiles/).  These
// versions assume what the kernels guarantee -- positive, finite, normal arguments -- and fall back
// to libdevice (out of line) otherwise.  Polynomial coefficients live in __constant__ memory, which
// the compiler fetches two at a time into uniform registers (LDCU.128).
//
// Accuracy (checked against libdevice / mpmath in tests): fast_log < 2e-16 absolute + 1 ulp relative,
// fast_rcp / fast_sqrt <= 1 ulp, sincos_u32 < 2e-16, lgamma_pos < 4e-16 relative (x >= 1e-300).
#pragma once
#include "common.cuh"

namespace bf {

// log1p(r) = r + r^2 (c0 + c1 r + ... + c5 r^5) on |r| <= 2^-8 (Chebyshev fit, error 4e-19)
// sin(a) = a S(a^2), cos(a) = C(a^2) on |a| <= pi/4 (Chebyshev fits, errors < 1e-16)
static __constant__ double FM_LOG[6] = {-0.5, 0.33333333333333337, -0.2499999999822793, 0.19999999998424825,
                                        -0.16666964331747255, 0.1428597887698295};
static __constant__ double FM_SIN[6] = {-0.16666666666666616, 0.008333333333320356, -0.0001984126982864526,
                                        2.7557313374649857e-06, -2.5050716696695676e-08, 1.5894720300223109e-10};
static __constant__ double FM_COS[6] = {0.04166666666666645, -0.0013888888888861082, 2.4801587283874084e-05,
                                        -2.755731309595641e-07, 2.0875582146477433e-09, -1.1353367957830792e-11};
// Stirling: 1/12, -1/360, 1/1260, -1/1680, 1/1188 ; then log(2 pi)/2, log 2
static __constant__ double FM_STI[8] = {8.333333333333333333e-2, -2.777777777777777778e-3, 7.936507936507936508e-4,
                                        -5.952380952380952381e-4, 8.417508417508417508e-4,
                                        0.918938533204672741780329736406, 0.693147180559945309417232121458, 0.0};

// log1p(t) = t (1 - t/2 + t^2/3 - ... - t^9/10) for |t| <= 1/32 (truncation < 3e-18)
static __constant__ double FM_L1P[10] = {-1.0 / 10, 1.0 / 9, -1.0 / 8, 1.0 / 7, -1.0 / 6, 1.0 / 5, -1.0 / 4, 1.0 / 3, -1.0 / 2, 1.0};
__device__ __forceinline__ double log1p_series(double t) {
  double s = fma(FM_L1P[0], t, FM_L1P[1]);
#pragma unroll
  for (int j = 2; j < 9; j++) s = fma(s, t, FM_L1P[j]);
  return fma(s * t, t, t);
}

constexpr int FM_LOG_TBL = 128;   // entries of the shared-memory table used by fast_log

// tbl[j] = (1/c_j, -log(1/c_j)) with c_j the midpoint of the j-th of 128 mantissa intervals of [1, 2).
// One statically allocated table per kernel (a fixed shared-memory address: no pointer to carry or
// recompute), built once per block (blocks are persistent); the caller synchronises before the first use.
__device__ __forceinline__ double2* fm_log_table() {
  __shared__ double2 tab[FM_LOG_TBL];
  return tab;
}
__device__ __forceinline__ void build_log_table() {
  double2* tbl = fm_log_table();
  for (int j = threadIdx.x; j < FM_LOG_TBL; j += blockDim.x) {
    const double inv = 1.0 / (1.0 + (j + 0.5) * (1.0 / FM_LOG_TBL));
    tbl[j] = make_double2(inv, -log(inv));
  }
}

// log(x), x positive finite normal (anything else: libdevice).  x = 2^e m, m in [1,2); r = m/c_j - 1,
// |r| <= 2^-8; log x = e log 2 + log c_j + log1p(r).
__device__ __forceinline__ double fast_log(double x) {
  const double2* tbl = fm_log_table();
  const int hi = __double2hiint(x);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return nl_log(x);
  const double2 t = tbl[(hi >> 13) & (FM_LOG_TBL - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, t.x, -1.0);
  const double e = (double)((hi >> 20) - 1023);
  double p = fma(r, FM_LOG[5], FM_LOG[4]);
  p = fma(p, r, FM_LOG[3]); p = fma(p, r, FM_LOG[2]); p = fma(p, r, FM_LOG[1]); p = fma(p, r, FM_LOG[0]);
  return fma(e, FM_STI[6], t.y) + fma(p * r, r, r);
}

// the same without the range check: the caller guarantees a positive, finite, normal argument (uniforms, ratios)
__device__ __forceinline__ double fast_log_pos(double x) {
  const double2* tbl = fm_log_table();
  const int hi = __double2hiint(x);
  const double2 t = tbl[(hi >> 13) & (FM_LOG_TBL - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, t.x, -1.0);
  const double e = (double)((hi >> 20) - 1023);
  double p = fma(r, FM_LOG[5], FM_LOG[4]);
  p = fma(p, r, FM_LOG[3]); p = fma(p, r, FM_LOG[2]); p = fma(p, r, FM_LOG[1]); p = fma(p, r, FM_LOG[0]);
  return fma(e, FM_STI[6], t.y) + fma(p * r, r, r);
}

// exp(x) for x <= 0 (0 below -700): x = k log 2 + r, |r| <= log(2)/2; exp(r) by its Taylor series to r^13 (remainder
// < 5e-18), scaled by 2^k through the exponent field.  Relative error < 3e-16.
static __constant__ double FM_EXP[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                         1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
__device__ __forceinline__ double fast_exp_nonpos(double x) {
  const double xc = x > -700.0 ? x : -700.0;
  const double t = fma(xc, 1.4426950408889634074, 6755399441055744.0);      // round(x / log 2) in the low word
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -0.693147180559945286, xc);
  r = fma(kf, -2.319046813846299558e-17, r);
  double p = fma(r, FM_EXP[0], FM_EXP[1]);
#pragma unroll
  for (int j = 2; j < 12; j++) p = fma(p, r, FM_EXP[j]);
  p = fma(p * r, r, r) + 1.0;                                                // 1 + r + r^2 (1/2 + ...)
  const double s = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  return x > -700.0 ? s : 0.0;
}

// 1/x and 1/sqrt(x), sqrt(x): MUFU seed (about 20 bits) + two Newton steps; x positive finite normal,
// result normal.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double fast_rcp1(double x) {             // one Newton step: about 40 bits
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return fma(y, fma(-x, y, 1.0), y);
}
__device__ __forceinline__ double fast_rsqrt_seed1(double x) {      // one Newton step: about 40 bits
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  const double t = fma(-h * y, y, 0.5);
  return fma(y, t, y);
}
__device__ __forceinline__ double fast_sqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double t = fma(-h * y, y, 0.5);
  y = fma(y, t, y);
  t = fma(-h * y, y, 0.5);
  y = fma(y, t, y);
  double g = x * y;
  const double r = fma(-g, g, x);
  return fma(0.5 * y, r, g);
}

// A uniformly distributed point (cs, sn) on the unit circle from one 32-bit word: bits 31/30 are the
// signs, bit 29 swaps the coordinates, the low 29 bits give the angle in (0, pi/4).
__device__ __forceinline__ void sincos_u32(uint32_t w, double& cs, double& sn) {
  const double al = fma((double)(w & 0x1fffffffu), 1.4629180792671596e-09, 7.314590396335798e-10);
  const double u = al * al;
  double s = fma(u, FM_SIN[5], FM_SIN[4]);
  s = fma(s, u, FM_SIN[3]); s = fma(s, u, FM_SIN[2]); s = fma(s, u, FM_SIN[1]); s = fma(s, u, FM_SIN[0]);
  s = fma(s * u, al, al);
  double c = fma(u, FM_COS[5], FM_COS[4]);
  c = fma(c, u, FM_COS[3]); c = fma(c, u, FM_COS[2]); c = fma(c, u, FM_COS[1]); c = fma(c, u, FM_COS[0]);
  c = fma(c * u, u, fma(u, -0.5, 1.0));
  const bool swap = (w >> 29) & 1u;
  const double a = swap ? s : c, b = swap ? c : s;
  cs = __hiloint2double(__double2hiint(a) ^ (int)(w & 0x80000000u), __double2loint(a));
  sn = __hiloint2double(__double2hiint(b) ^ (int)((w << 1) & 0x80000000u), __double2loint(b));
}

// (0,1) uniform with 52 random bits: mantissa trick, no integer -> double conversion
__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo) {
  const double d = __hiloint2double((int)(0x3ff00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
  return d - (1.0 - 1.1102230246251565e-16);      // [2^-53, 1 - 2^-53]
}

// two independent standard normals from three words (Box-Muller)
__device__ __forceinline__ void fast_box_muller(uint32_t w0, uint32_t w1, uint32_t w2, double& n0, double& n1) {
  const double r = fast_sqrt(-2.0 * fast_log_pos(u52(w0, w1)));      // u52 is in [2^-53, 1): positive and normal
  double cs, sn;
  sincos_u32(w2, cs, sn);
  n0 = r * cs; n1 = r * sn;
}

// Stirling's series for x >= 16 given lx = log x (exact to rounding: the next term is < 1.1e-16)
__device__ __forceinline__ double stirling16(double x, double lx) {
  const double r = fast_rcp(x), r2 = r * r;
  double s = fma(r2, FM_STI[4], FM_STI[3]);
  s = fma(r2, s, FM_STI[2]); s = fma(r2, s, FM_STI[1]); s = fma(r2, s, FM_STI[0]);
  return fma(x - 0.5, lx, -x) + fma(r, s, FM_STI[5]);
}
// S(r) = lgamma(x) - [(x - 1/2) log x - x + log(2 pi)/2] at r = 1/x, x >= 16 (the series' next term is < 1.1e-16)
__device__ __forceinline__ double stirling_corr(double r) {
  const double r2 = r * r;
  double s = fma(r2, FM_STI[4], FM_STI[3]);
  s = fma(r2, s, FM_STI[2]); s = fma(r2, s, FM_STI[1]); s = fma(r2, s, FM_STI[0]);
  return r * s;
}
// log Gamma(x), x > 0, with lx = log x known.  x < 16 is shifted: Gamma(x) = Gamma(x + 16) / (x (x+1) ... (x+15));
// that branch is out of line (one copy per kernel instead of one per call site: the Z kernel is
// instruction-cache sensitive).
static __device__ __noinline__ double lgamma_shift16(double x, double lx) {
  double pr = (x + 1.0) * (x + 2.0);
#pragma unroll
  for (int j = 3; j < 15; j += 2) pr *= fma(x, x + (2 * j + 1), (double)(j * (j + 1)));   // (x+j)(x+j+1)
  pr *= (x + 15.0);
  const double xs = x + 16.0;
  return stirling16(xs, fast_log(xs)) - lx - fast_log(pr);
}
__device__ __forceinline__ double lgamma_pos(double x, double lx) {
  if (x >= 16.0) return stirling16(x, lx);
  return lgamma_shift16(x, lx);
}
static __device__ __noinline__ double fast_log_nl(double x) { return fast_log(x); }

// One Marsaglia-Tsang candidate for shape >= 1 from a given normal x and accept-uniform uu.  Returns the
// candidate g = d v and whether it is accepted (false with probability ~1e-3 at shape 10, ~1e-5 at
// shape 3000: the caller then falls back to RngStream::gamma).  c = 1/sqrt(9d) only shapes the envelope
// (any c gives an exact sampler as long as the same c is used in the test), so a 40-bit value is enough.
// l3 receives 3 log(1 + c x) = log v (the Z step reuses it: log g = log d + l3).
__device__ __forceinline__ bool gamma_candidate_fast(double shape, double x, double uu, double& g, double& l3) {
  const double d = shape - 1.0 / 3.0;
  const double c = fast_rsqrt_seed1(9.0 * d);
  const double t = c * x;
  const double v1 = 1.0 + t;
  const double v = v1 * v1 * v1;
  g = d * v;
  l3 = 0.0;
  if (t <= -0.99) return false;
  l3 = 3.0 * (fabs(t) <= 0.03125 ? log1p_series(t) : fast_log_nl(v1));
  const double R = fma(0.5 * x, x, d * (1.0 - v + l3));
  bool ok = (uu - 1.0 < R);
  if (!ok) ok = fast_log_nl(uu) < R;
  return ok;
}

}  // namespace bf
