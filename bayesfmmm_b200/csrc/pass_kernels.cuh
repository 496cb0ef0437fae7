// pass_kernels.cuh -- the three per-function passes (Z Metropolis step, chi sweep, residual
// sum of squares) over the projected coefficient cache.  One thread owns V (1 or 2) adjacent
// functions; every global access is a 16-byte load/store on coefficient-major (SoA) arrays, so a
// warp reads 512 contiguous bytes per row.  The global coefficients live in shared memory
// (broadcast reads).  K and M are compile-time so all per-function state stays in registers.
#pragma once
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "fastmath.cuh"

namespace bf {

// V functions per thread: V = 2 -> 16-byte accesses, V = 1 -> 8-byte accesses (half the registers,
// twice the resident warps: used by the latency-bound kernels).
template <int V> __device__ __forceinline__ void ldv(const double* p, double (&o)[V]) {
  if constexpr (V == 2) { double2 t = *reinterpret_cast<const double2*>(p); o[0] = t.x; o[1] = t.y; }
  else o[0] = *p;
}
template <int V> __device__ __forceinline__ void ldv_cs(const double* p, double (&o)[V]) {
  if constexpr (V == 2) { double2 t = __ldcs(reinterpret_cast<const double2*>(p)); o[0] = t.x; o[1] = t.y; }
  else o[0] = __ldcs(p);
}
template <int V> __device__ __forceinline__ void stv(double* p, const double (&o)[V]) {
  if constexpr (V == 2) *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
  else *p = o[0];
}

// Streams the P rows of the coefficient cache for the V functions of this thread with the loads of
// the NEXT group of CH rows in flight while the current group is consumed (software pipelining:
// without it the passes sit on long-scoreboard stalls, see profiles/).  Loads are ld.global.cs
// (streamed once per pass, evict-first).  NB > 0 (ragged grids) also streams, for every row p, the
// bw <= NB band rows Gl[j*P + p] of the per-function Gram matrices.
template <int V, int CH, int NB, typename F>
__device__ __forceinline__ void stream_rows(const double* __restrict__ base, const double* __restrict__ band, int bw,
                                            int ld, int P, int i0, F&& body) {
  double cur[CH][V], nxt[CH][V];
  double gcur[CH][NB > 0 ? NB : 1][V], gnxt[CH][NB > 0 ? NB : 1][V];
  const double* col = base + i0;
  const double* bcol = band + i0;
  auto fetch = [&](int p, double (&c)[V], double (&g)[NB > 0 ? NB : 1][V]) {
    if (p < P) {
      ldv_cs<V>(col + (size_t)p * ld, c);
      if constexpr (NB > 0) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
          if (j < bw) ldv_cs<V>(bcol + ((size_t)j * P + p) * ld, g[j]);
          else {
#pragma unroll
            for (int v = 0; v < V; v++) g[j][v] = 0.0;
          }
        }
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; v++) c[v] = 0.0;
    }
  };
#pragma unroll
  for (int j = 0; j < CH; j++) fetch(j, cur[j], gcur[j]);
  for (int p0 = 0; p0 < P; p0 += CH) {
#pragma unroll
    for (int j = 0; j < CH; j++) fetch(p0 + CH + j, nxt[j], gnxt[j]);
#pragma unroll
    for (int j = 0; j < CH; j++)
      if (p0 + j < P) body(p0 + j, cur[j], gcur[j]);
#pragma unroll
    for (int j = 0; j < CH; j++) {
#pragma unroll
      for (int v = 0; v < V; v++) cur[j][v] = nxt[j][v];
      if constexpr (NB > 0) {
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
          for (int v = 0; v < V; v++) gcur[j][b][v] = gnxt[j][b][v];
      }
    }
  }
}

// Sliding window for the banded quadratic / bilinear forms of the ragged-grid kernels.
//   x' G y = sum_p [ G[p][p] x_p y_p + sum_{j>=1} G[p-j][p] (x_{p-j} y_p + x_p y_{p-j}) ]
// prev[j-1] holds the value of the vector at row p - j.
struct BandWin {
  double prev[BWMAX - 1];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < BWMAX - 1; j++) prev[j] = 0.0;
  }
  __device__ __forceinline__ void push(double x) {
#pragma unroll
    for (int j = BWMAX - 2; j > 0; j--) prev[j] = prev[j - 1];
    prev[0] = x;
  }
};
// contribution of row p to x' G y (x, y the current values; wx, wy the windows BEFORE pushing p)
__device__ __forceinline__ double band_term(const double (&g)[BWMAX], double x, double y, const BandWin& wx,
                                            const BandWin& wy) {
  double t = g[0] * x * y;
#pragma unroll
  for (int j = 1; j < BWMAX; j++) t = fma(g[j], fma(wx.prev[j - 1], y, x * wy.prev[j - 1]), t);
  return t;
}
// x' G x
__device__ __forceinline__ double band_term_sq(const double (&g)[BWMAX], double x, const BandWin& wx) {
  double t = g[0] * x;
#pragma unroll
  for (int j = 1; j < BWMAX; j++) t = fma(2.0 * g[j], wx.prev[j - 1], t);
  return t * x;
}
constexpr int ROW_CH_RAGGED = 2; // ragged grids: every row also brings bw band rows

// Effective coefficients of one function at basis column p:
//   a[k][0]   = (nu_k + eta_k x_i)[p],   a[k][m+1] = (phi_km + xi_km x_i)[p]     (whitened)
template <int K, int M, bool COV, int V>
struct Coef {
  static constexpr int NV = COV ? V : 1;
  double a[NV][K][M + 1];
  __device__ __forceinline__ void load(const double* __restrict__ gs, int D, const double (&x)[V][DMAX]) {
    if constexpr (!COV) {
      constexpr int Q = K * (M + 1);
      double flat[Q + 1];
#pragma unroll
      for (int j = 0; j < (Q + 1) / 2; j++) {
        double2 t = *reinterpret_cast<const double2*>(gs + 2 * j);
        flat[2 * j] = t.x;
        if (2 * j + 1 < Q + 1) flat[2 * j + 1] = t.y;
      }
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int m = 0; m <= M; m++) a[0][k][m] = flat[k * (M + 1) + m];
    } else {
      const int stride = 1 + D;
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int m = 0; m <= M; m++) {
          const double* b = gs + (k * (M + 1) + m) * stride;
          double base = b[0];
#pragma unroll
          for (int v = 0; v < V; v++) a[v][k][m] = base;
#pragma unroll
          for (int d = 0; d < DMAX; d++)
            if (d < D) {
              double g = b[1 + d];
#pragma unroll
              for (int v = 0; v < V; v++) a[v][k][m] = fma(x[v][d], g, a[v][k][m]);
            }
        }
    }
  }
  __device__ __forceinline__ double get(int v, int k, int m) const { return a[COV ? v : 0][k][m]; }
};

template <int K, int M, bool COV, int V>
struct FnState {
  double z[V][K], chi[V][M > 0 ? M : 1], x[V][DMAX];
  __device__ __forceinline__ void load(const PassArgs& a, int i0) {
    double t[V];
#pragma unroll
    for (int k = 0; k < K; k++) {
      ldv<V>(a.Z + (size_t)k * a.ld + i0, t);
#pragma unroll
      for (int v = 0; v < V; v++) z[v][k] = t[v];
    }
#pragma unroll
    for (int m = 0; m < M; m++) {
      ldv<V>(a.chi + (size_t)m * a.ld + i0, t);
#pragma unroll
      for (int v = 0; v < V; v++) chi[v][m] = t[v];
    }
#pragma unroll
    for (int d = 0; d < DMAX; d++)
#pragma unroll
      for (int v = 0; v < V; v++) x[v][d] = 0;
    if constexpr (COV) {
#pragma unroll
      for (int d = 0; d < DMAX; d++)
        if (d < a.D) {
          ldv<V>(a.X + (size_t)d * a.ld + i0, t);
#pragma unroll
          for (int v = 0; v < V; v++) x[v][d] = t[v];
        }
    }
  }
};

__device__ __forceinline__ void stage_globals(const PassArgs& a, double* g) {
  const int tot = a.P4 * a.QS;
  for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) g[idx] = a.glob[idx];
  __syncthreads();
}

// ================================================================= Z Metropolis step
// updateZ_PM and twins (reference UpdateMixedMembership.h:131-185; lpdf_z :20-50; tempered :91;
// Z_proposal_density :102-113; rdirichlet Distributions.h:22-45; calc_lB :51-61).
// The squared-error terms are evaluated in the whitened coefficient space, where
// ||y - B theta||^2 = rss_i + ||c~_i - theta~||^2 and rss_i cancels in the ratio.
// No logarithm of a membership and no log-Gamma is evaluated on the common path (closed forms below);
// the remaining transcendentals go through fastmath.cuh.

// Rows of the coefficient cache in groups of four, two groups alternating (A is consumed while B is in
// flight and vice versa: no register copies).  begin() issues the first group, so the caller can put
// independent work (the proposal) between it and run().  The cache and the globals are padded with
// zero rows up to P4 = a multiple of four (they add nothing to any sum), so there are no predicates.
template <int V>
struct RowStream {
  double A[4][V], B[4][V];
  const double* q;
  size_t s1;
  __device__ __forceinline__ void fetch(const double* r, double (&c)[4][V]) {
#pragma unroll
    for (int j = 0; j < 4; j++) ldv_cs<V>(r + j * s1, c[j]);
  }
  __device__ __forceinline__ void begin(const double* col, int ld) {
    q = col; s1 = (size_t)ld;
    fetch(q, A);
  }
  template <typename F>
  __device__ __forceinline__ void run(int P4, F&& body) {
    for (int p = 0; p < P4; p += 8) {
      const bool second = p + 4 < P4;
      if (second) fetch(q + 4 * s1, B);
#pragma unroll
      for (int j = 0; j < 4; j++) body(p + j, A[j]);
      q += 8 * s1;
      if (p + 8 < P4) fetch(q, A);
      if (second) {
#pragma unroll
        for (int j = 0; j < 4; j++) body(p + 4 + j, B[j]);
      }
    }
  }
};

// Generic proposal (K > 4, or a membership that is not positive: the reference's rdirichlet replaces its
// concentration by 10, Distributions.h:24-28): the general gamma sampler on its own stream.  Out of line, rare.
template <int K>
__device__ __noinline__ void z_slow_proposal(uint64_t key, uint64_t gi, uint64_t iteration, const double* z, double a_Z, double* zp,
                                             double* uacc, bool live) {
  RngStream rs(key, gi, iteration, RNG_Z_PROPOSAL_SLOW);
#pragma unroll
  for (int k = 0; k < K; k++) {
    double sh = a_Z * z[k];
    if (!(sh > 0)) sh = 10;                                              // Distributions.h:24-28
    zp[k] = live ? rs.gamma(sh) : 1.0;
  }
  if (K > 4) *uacc = rs.uniform();
}

// The Metropolis log-ratio without the likelihood term, written out as the reference does
// (UpdateMixedMembership.h:33-36,102-113,165-168; Distributions.h:51-61):
//   sum_k (alpha_3 pi_k - 1)(log z*_k - log z_k) + q(z | a z*) - q(z* | a z),
//   q(x | al) = sum_k (al_k - 1) log x_k - [sum_k lgamma(al_k) - lgamma(sum_k al_k)].
// Out of line: only functions outside the domain of the closed form below come here (a membership that is not
// positive, a proposal coordinate that underflowed to zero, rows of Z that do not sum to one).
template <int K>
__device__ __noinline__ double z_logratio_exact(const double* z, const double* zp, const double* pi, double alpha3, double a) {
  double lp = 0, q_new = 0, q_old = 0, lB_new = 0, lB_old = 0, tot_new = 0, tot_old = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double lz = log(z[k]), lzp = log(zp[k]);
    const double al_from_old = a * z[k], al_from_new = a * zp[k];
    lp = fma(alpha3 * pi[k] - 1, lzp - lz, lp);
    q_new = fma(al_from_old - 1, lzp, q_new);
    q_old = fma(al_from_new - 1, lz, q_old);
    lB_new += lgamma(al_from_old); tot_new += al_from_old;
    lB_old += lgamma(al_from_new); tot_old += al_from_new;
  }
  q_new -= (lB_new - lgamma(tot_new));
  q_old -= (lB_old - lgamma(tot_old));
  return lp + (q_old - q_new);
}

// S(1/x) = lgamma(x) - [(x - 1/2) log x - x + log(2 pi)/2] for any x > 0 below 16, where Stirling's series is not
// accurate: through the shift Gamma(x) = Gamma(x + 16) / (x ... (x + 15)).  S(1/xp) - S(1/x) for a coordinate whose
// current or proposed concentration is below 16 (a membership below 16 / a): out of line, one call per such coordinate.
static __device__ __noinline__ double stirling_corr_diff_small(double x, double rx, double xp, double rxp) {
  double so = stirling_corr(rx), sn = stirling_corr(rxp);
  if (x < 16.0) { const double lx = fast_log(x); so = lgamma_shift16(x, lx) - (fma(x - 0.5, lx, -x) + 0.918938533204672741780329736406); }
  if (xp < 16.0) { const double lx = fast_log(xp); sn = lgamma_shift16(xp, lx) - (fma(xp - 0.5, lx, -xp) + 0.918938533204672741780329736406); }
  return sn - so;
}

#ifndef BF_Z_MINB
#define BF_Z_MINB 6       // resident blocks per SM the register allocation of the V = 1 common-grid Z kernel targets
#endif
// The log-ratio in closed form.  With sh_k = a z_k, sh*_k = a z*_k, dl_k = log z*_k - log z_k and Stirling's
// lgamma(x) = (x - 1/2) log x - x + log(2 pi)/2 + S(1/x), every log-Gamma main term folds into the log terms:
//   q(z | a z*) - q(z* | a z) = - sum_k (sh_k + sh*_k - 3/2) dl_k - sum_k [S(1/sh*_k) - S(1/sh_k)]
//                               + (1 - log a)(tot* - tot) + lgamma(tot*) - lgamma(tot),     tot = sum_k sh_k,
// so no log-Gamma is evaluated and the only logarithm per coordinate is the one of the ratio z*_k / z_k: the summands
// are O(a |dl|) instead of the O(a log a) of the 2(K + 1) log-Gammas the reference sums (its own rounding error,
// 2(K+1) ulp(lgamma(a)) ~ 1e-10 at a = 1e4, is what bounds the parity of this quantity; see tests/test_gpu_parity.py).
// The identity holds for every positive z, z*; S itself is Stirling's series above 16 and the shifted form below.
// lgamma(tot*) - lgamma(tot) is the second-order Taylor expansion around a (rows of Z sum to one up to rounding;
// `ok` is cleared otherwise and the caller evaluates the reference's formula as written).
template <int K>
__device__ __forceinline__ double z_logratio_closed(const double (&sh)[K], const double (&r)[K], const double (&zp)[K],
                                                    const double (&dl)[K], const PassArgs& a, const double* par, bool& ok) {
  double acc = 0, tot = 0, tots = 0;
  const double alpha3 = par[8];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double shp = a.a_Z_PM * zp[k];
    const double rp = fast_rcp1(shp);
    acc = fma(fma(alpha3, par[k], 0.5) - (sh[k] + shp), dl[k], acc);
    double ds = stirling_corr(rp) - stirling_corr(r[k]);
    if (shp < 16.0 || sh[k] < 16.0) ds = stirling_corr_diff_small(sh[k], r[k], shp, rp);
    acc -= ds;
    tot += sh[k]; tots += shp;
  }
  const double dt = tot - a.a_Z_PM, dts = tots - a.a_Z_PM;
  ok = ok && (fabs(dt) <= 1e-8 * a.a_Z_PM) && (fabs(dts) <= 1e-8 * a.a_Z_PM);
  acc = fma(dts - dt, a.c_tot, acc);                       // c_tot = 1 - log a + digamma(a)
  return fma(0.5 * a.trigam_a, (dts - dt) * (dts + dt), acc);
}

// Marsaglia-Tsang's acceptance test as written: log uu < x^2/2 + d (1 - v + log v), v = v1^3 (out of line: the
// squeeze in z_candidate_round decides almost every candidate)
static __device__ __noinline__ bool mt_accept_exact(double v1, double vv, double d, double x, double uu) {
  const double l3 = 3.0 * fast_log_pos(v1);
  const double R = fma(0.5 * x, x, d * (1.0 - vv + l3));
  return (uu - 1.0 < R) || (fast_log_nl(uu) < R);
}

// One round of Marsaglia-Tsang candidates for the coordinates still pending (bit k of `pend`): three Philox blocks
// give two Box-Muller pairs (4 normals) and four 32-bit accept uniforms.  Candidate g = d v, d = s - 1/3,
// v = (1 + c x)^3 with s = sh (sh >= 1) or sh + 1 (boosted below); accepted when log uu < x^2/2 + d (1 - v + log v),
// decided without a logarithm when uu - 1 < R (log uu <= uu - 1).  Coordinates are independent rejection samplers, so
// a rejected one simply draws again in the next round (blocks 3 r .. 3 r + 2); at the default a_Z_PM a round rejects
// ~1e-5 of the candidates, ~5e-2 of those with sh < 2.
template <int K>
__device__ __forceinline__ unsigned z_candidate_round(const PassArgs& a, uint64_t gi, uint32_t block0, const double (&sh)[K],
                                                      unsigned pend, double (&g)[K], double& uacc, bool first) {
  uint32_t w0[4], w1[4], w2[4];
  philox_rk(a, gi, RNG_Z_PROPOSAL, block0, w0);
  philox_rk(a, gi, RNG_Z_PROPOSAL, block0 + 1, w1);
  philox_rk(a, gi, RNG_Z_PROPOSAL, block0 + 2, w2);
  double nrm[4] = {0, 0, 0, 0}, ua[4];
  fast_box_muller(w0[0], w0[1], w0[2], nrm[0], nrm[1]);
  if constexpr (K > 2) fast_box_muller(w1[0], w1[1], w1[2], nrm[2], nrm[3]);
  ua[0] = u32(w0[3]); ua[1] = u32(w1[3]); ua[2] = u32(w2[2]); ua[3] = u32(w2[3]);
  if (first) uacc = u52(w2[0], w2[1]);
  unsigned left = pend;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double d = (sh[k] < 1.0 ? sh[k] + 1.0 : sh[k]) - 1.0 / 3.0;
    const double c = fast_rsqrt_seed1(9.0 * d);
    const double t = c * nrm[k];
    const double v1 = 1.0 + t;
    const double vv = v1 * v1 * v1;
    const bool in = t > -0.99;
    if (((pend >> k) & 1u) != 0) {
      g[k] = d * vv;
      // R = x^2/2 + d (1 - v + log v) = 3 d sum_{j >= 4} (-1)^(j+1) t^j / j  >=  -(3/4) d t^4 (1 + 2|t|)  for |t| <= 1/2:
      // a candidate with log uu <= uu - 1 below that bound is accepted without a logarithm (all but ~2e-4 of them at
      // the default a_Z_PM); the rest take the exact test out of line
      const double t2 = t * t;
      bool ok = (fabs(t) <= 0.5) && (ua[k] - 1.0 < -0.75 * d * t2 * t2 * fma(2.0, fabs(t), 1.0));
      if (in && !ok) ok = mt_accept_exact(v1, vv, d, nrm[k], ua[k]);
      if (ok) left &= ~(1u << k);
    }
  }
  return left;
}

// The proposal of one function and everything of the Metropolis ratio that does not need the coefficient cache:
// z* ~ Dirichlet(a z) normalised (rdirichlet, Distributions.h:22-45), lr = log-ratio of prior and proposal densities,
// lu = log of the Metropolis uniform.  Reads Z only -- not the globals, chi or sigma^2.
template <int K>
__device__ __forceinline__ void z_propose(const PassArgs& a, const double* s_par, const double (&z)[K], int idx,
                                          double (&zp)[K], double& lr, double& lu) {
  const uint64_t gi = a.global_offset + (uint64_t)idx;
  const bool live = idx < a.n;
  double sh[K], rr[K], dl[K], uacc = 0.5, sum = 0;
  bool generic = (K > 4);        // the gamma variates must come from the generic sampler
  bool pos = true;               // every membership is positive (and normal): the closed forms apply
  if (a.gam) {                   // injected draws (parity)
#pragma unroll
    for (int k = 0; k < K; k++) {
      sh[k] = a.a_Z_PM * z[k];
      pos = pos && (sh[k] > 1e-290);
      zp[k] = __ldcs(a.gam + (size_t)k * a.ld + idx); sum += zp[k];
    }
    uacc = __ldcs(a.u + idx);
    generic = false;
  } else if constexpr (K <= 4) {
    // the first round's random words and normals do not depend on the state: the loads issued above are in
    // flight while they are made
    double gk[K];
#pragma unroll
    for (int k = 0; k < K; k++) sh[k] = 1.0;
    unsigned pend = (1u << K) - 1u;
    // (sh is filled in below, after the words of round 0 are under way: z_candidate_round reads it late)
    bool small = false;
#pragma unroll
    for (int k = 0; k < K; k++) {
      sh[k] = a.a_Z_PM * (live ? z[k] : 1.0 / K);
      pos = pos && (sh[k] > 1e-290);                                  // false for NaN and z <= 0 too
      small = small || (sh[k] < 1.0);
    }
    generic = !pos;
    pend = z_candidate_round<K>(a, gi, 0, sh, pend, gk, uacc, true);
    for (uint32_t round = 1; round < 24 && __any_sync(__activemask(), pend != 0); round++)
      pend = z_candidate_round<K>(a, gi, 8 + 3 * round, sh, pend, gk, uacc, false);
    generic = generic || (pend != 0);
    // shapes below 1 (tiny memberships): Gamma(s) = Gamma(s + 1) U^(1/s); the extra uniforms are only generated by
    // warps that contain such a function
    if (__any_sync(__activemask(), small)) {
      uint32_t w3[4], w4[4];
      philox_rk(a, gi, RNG_Z_PROPOSAL, 3, w3);
      philox_rk(a, gi, RNG_Z_PROPOSAL, 4, w4);
      const double ub[4] = {u52(w3[0], w3[1]), u52(w3[2], w3[3]), u52(w4[0], w4[1]), u52(w4[2], w4[3])};
#pragma unroll
      for (int k = 0; k < K; k++)
        if (sh[k] < 1.0) gk[k] *= fast_exp_nonpos(fast_log_pos(ub[k]) * fast_rcp(pos ? sh[k] : 1.0));
    }
#pragma unroll
    for (int k = 0; k < K; k++) { zp[k] = gk[k]; sum += gk[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < K; k++) { sh[k] = a.a_Z_PM * z[k]; pos = pos && (sh[k] > 1e-290); }
  }
  if (generic) {
    double tz[K], tzp[K], tu = uacc;     // only these copies have their address taken
#pragma unroll
    for (int k = 0; k < K; k++) tz[k] = z[k];
    z_slow_proposal<K>(a.key, gi, a.iteration, tz, a.a_Z_PM, tzp, &tu, live);
    sum = 0;
#pragma unroll
    for (int k = 0; k < K; k++) { zp[k] = tzp[k]; sum += tzp[k]; }
    uacc = tu;
  }
  if (a.draws_out) {
#pragma unroll
    for (int k = 0; k < K; k++) a.draws_out[(size_t)k * a.ld + idx] = zp[k];
    a.draws_out[(size_t)K * a.ld + idx] = uacc;
  }
  if (a.gam || generic) {          // the reference's division (Distributions.h:39-43): same bits as the oracle
#pragma unroll
    for (int k = 0; k < K; k++) zp[k] = zp[k] / sum;
  } else {                         // straight-line path: the sum of K positive normal numbers
    const double rs = fast_rcp(sum);
#pragma unroll
    for (int k = 0; k < K; k++) zp[k] *= rs;
  }
  // ---- log-ratio of prior and proposal densities
  bool closed = pos;
#pragma unroll
  for (int k = 0; k < K; k++) closed = closed && (zp[k] > 1e-290);
#pragma unroll
  for (int k = 0; k < K; k++) {
    rr[k] = fast_rcp(closed ? sh[k] : 1.0);
    dl[k] = fast_log_pos(closed ? zp[k] * (a.a_Z_PM * rr[k]) : 1.0);      // log(z*_k / z_k)
  }
  lr = z_logratio_closed<K>(sh, rr, zp, dl, a, s_par, closed);
  if (!closed) {
    double tz[K], tzp[K], tpi[K];
#pragma unroll
    for (int k = 0; k < K; k++) { tz[k] = z[k]; tzp[k] = zp[k]; tpi[k] = s_par[k]; }
    lr = z_logratio_exact<K>(tz, tzp, tpi, s_par[8], a.a_Z_PM);
  }
  lu = fast_log(uacc);
}

template <int K, int M, bool COV, int V, bool RG, bool PRE = false>
__global__ void __launch_bounds__(PF_THREADS, (V == 1 && !RG) ? (PRE ? 8 : BF_Z_MINB) : 4) z_kernel(const PassArgs a) {
  extern __shared__ double g[];
  // pi, alpha_3, sigma^2: kernel arguments, or device memory when the sweep's small updates run on the device
  __shared__ double s_par[10];
  if (threadIdx.x < 10)
    s_par[threadIdx.x] = a.zpar_dev ? a.zpar_dev[threadIdx.x] : (threadIdx.x < 8 ? a.pi[threadIdx.x] : (threadIdx.x == 8 ? a.alpha3 : a.sigma_sq));
  build_log_table();
  stage_globals(a, g);
  double red[K + 1];
#pragma unroll
  for (int j = 0; j <= K; j++) red[j] = 0;
  const double hb = a.beta / (2 * s_par[9]);
  // persistent blocks: every thread walks the functions with a grid stride, one reduction per block
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    FnState<K, M, COV, V> st;
    st.load(a, i0);
    RowStream<V> rows;
    if constexpr (!RG) rows.begin(a.Ct + i0, a.ld);      // first rows in flight during the proposal
    // Everything that does not need the coefficient cache comes first -- the proposal z*, the Metropolis uniform and
    // the log-ratio lr of prior and proposal densities -- so that only z, z*, chi, lr and log u are live across the
    // row loop (the loop's own state is what sets the register count, not the transcendental code).
    double zp[V][K], lr[V], lu[V];
    if constexpr (PRE) {       // made earlier by z_propose_kernel
      double t[V];
#pragma unroll
      for (int k = 0; k < K; k++) {
        ldv_cs<V>(a.zprop + (size_t)k * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) zp[v][k] = t[v];
      }
      ldv_cs<V>(a.zprop + (size_t)K * a.ld + i0, lr);
      ldv_cs<V>(a.zprop + (size_t)(K + 1) * a.ld + i0, lu);
    } else {
#pragma unroll
      for (int v = 0; v < V; v++) z_propose<K>(a, s_par, st.z[v], i0 + v, zp[v], lr[v], lu[v]);
    }
    // ---- squared errors of the current and the proposed state
    double so[V], sn[V];
#pragma unroll
    for (int v = 0; v < V; v++) { so[v] = 0; sn[v] = 0; }
    Coef<K, M, COV, V> cf;
    if constexpr (!RG) {
      const int QSc = COV ? a.QS : ((K * (M + 1) + 1) & ~1);      // compile-time without covariates
      auto row = [&](int p, const double (&c)[V]) {
        cf.load(g + p * QSc, a.D, st.x);
#pragma unroll
        for (int v = 0; v < V; v++) {
          double ro = c[v], rn = c[v];
#pragma unroll
          for (int k = 0; k < K; k++) {
            double at = cf.get(v, k, 0);
#pragma unroll
            for (int m = 0; m < M; m++) at = fma(st.chi[v][m], cf.get(v, k, m + 1), at);
            ro = fma(-st.z[v][k], at, ro);
            rn = fma(-zp[v][k], at, rn);
          }
          so[v] = fma(ro, ro, so[v]);
          sn[v] = fma(rn, rn, sn[v]);
        }
      };
      rows.run(a.P4, row);
    } else {
      constexpr int NB = BWMAX;
      BandWin wo[V], wn[V];
#pragma unroll
      for (int v = 0; v < V; v++) { wo[v].clear(); wn[v].clear(); }
      stream_rows<V, ROW_CH_RAGGED, NB>(a.Ct, a.Gl, a.bw, a.ld, a.P, i0,
                                        [&](int p, const double (&c)[V], const double (&gb)[NB][V]) {
        cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
        for (int v = 0; v < V; v++) {
          double ro = c[v], rn = c[v];
#pragma unroll
          for (int k = 0; k < K; k++) {
            double at = cf.get(v, k, 0);
#pragma unroll
            for (int m = 0; m < M; m++) at = fma(st.chi[v][m], cf.get(v, k, m + 1), at);
            ro = fma(-st.z[v][k], at, ro);
            rn = fma(-zp[v][k], at, rn);
          }
          // (c - theta)' G_i (c - theta) through the band of G_i
          double gv[BWMAX];
#pragma unroll
          for (int j = 0; j < BWMAX; j++) gv[j] = gb[j][v];
          so[v] += band_term_sq(gv, ro, wo[v]);
          sn[v] += band_term_sq(gv, rn, wn[v]);
          wo[v].push(ro); wn[v].push(rn);
        }
      });
    }
    // ---- acceptance
    double znew[V][K];
#pragma unroll
    for (int v = 0; v < V; v++) {
      bool nonpos = false;
#pragma unroll
      for (int k = 0; k < K; k++) nonpos |= (st.z[v][k] <= 0);
      double acc = fma(-hb, sn[v] - so[v], lr[v]);
      if (nonpos) acc = 1;                           // UpdateMixedMembership.h:170-174
      const bool live = (i0 + v) < a.n;
      const bool take = live && (lu[v] < acc);
      if (a.acc_out && live) a.acc_out[i0 + v] = acc;
#pragma unroll
      for (int k = 0; k < K; k++) {
        znew[v][k] = take ? zp[v][k] : st.z[v][k];
        if (live) red[k] += fast_log(znew[v][k]);    // sum_i log Z_ik of the new state: updatePi_PM / updateAlpha3's statistic
      }
      if (take) red[K] += 1.0;
    }
    {
      double t[V];
#pragma unroll
      for (int k = 0; k < K; k++) {
#pragma unroll
        for (int v = 0; v < V; v++) t[v] = znew[v][k];
        stv<V>(a.Z + (size_t)k * a.ld + i0, t);
      }
    }
  }
  grid_reduce<K + 1>(red, a);
}

// The proposal half of the Z step on its own (z_propose above): reads Z, writes (z*[K], lr, lu) per function.  It needs
// neither the globals nor chi nor sigma^2, so the sampler launches it for the NEXT sweep on a side stream while the host
// draws this sweep's Gaussian blocks; z_kernel<..., PRE = true> then only streams the cache and accepts.
#ifndef BF_ZP_MINB
#define BF_ZP_MINB 6
#endif
template <int K>
__global__ void __launch_bounds__(PF_THREADS, BF_ZP_MINB) z_propose_kernel(const PassArgs a) {
  __shared__ double s_par[10];
  if (threadIdx.x < 10)
    s_par[threadIdx.x] = a.zpar_dev ? a.zpar_dev[threadIdx.x] : (threadIdx.x < 8 ? a.pi[threadIdx.x] : (threadIdx.x == 8 ? a.alpha3 : a.sigma_sq));
  build_log_table();
  __syncthreads();
  for (int i = blockIdx.x * PF_THREADS + threadIdx.x; i < a.ld; i += gridDim.x * PF_THREADS) {
    double z[K], zp[K], lr, lu;
#pragma unroll
    for (int k = 0; k < K; k++) z[k] = a.Z[(size_t)k * a.ld + i];
    z_propose<K>(a, s_par, z, i, zp, lr, lu);
#pragma unroll
    for (int k = 0; k < K; k++) a.zprop_out[(size_t)k * a.ld + i] = zp[k];
    a.zprop_out[(size_t)K * a.ld + i] = lr;
    a.zprop_out[(size_t)(K + 1) * a.ld + i] = lu;
  }
}

// ================================================================= chi sweep (+ post-update SSR)
// updateChi and twins (reference UpdateChi.h:19-64; tempered :116-119).  Per function the M x M
// Gram G[m][n] = ph_m . ph_n and r[m] = ph_m . (y - B mu) are accumulated once (in coefficient
// space), then the reference's sequential m = 0..M-1 sweep is run on them, so chi(i,n) for n < m
// is the already-updated value exactly as in UpdateChi.h:48.
// CPO = true turns the same accumulation into the per-function MARGINAL log-likelihood (chi integrated
// out) that the reference's calcLikelihoodCPO evaluates per stored iteration (CalculateLikelihood.h:344-385):
// with U = B [u_1 .. u_M], cov = U U' + sigma^2 I and G = U'U (M x M, accumulated below), r = U'(y - B mu),
//   log det cov = (n_i - M) log sigma^2 + log det(sigma^2 I_M + G)
//   (y - B mu)' cov^{-1} (y - B mu) = (|y - B mu|^2 - r'(sigma^2 I_M + G)^{-1} r) / sigma^2
// (matrix determinant lemma / Woodbury), so the n_i x n_i covariance the reference factorises never exists.
// The CPO's harmonic mean over iterations is kept as a running log-sum-exp of -logl per function.
// Common basis without covariates: u_m = sum_k z_k phi~_km, so the per-function Gram is a quadratic form in z,
//   G_i[m][n] = u_m . u_n = sum_{k <= k'} z_k z_k' Q[mn][kk'],   Q[mn][kk'] = phi~_km . phi~_k'n (+ phi~_k'm . phi~_kn, k != k'),
// with Q a property of the globals only: every block builds it once from the staged globals.
template <int K, int M>
__device__ __forceinline__ void build_gram_forms(const PassArgs& a, const double* g, double* Qs) {
  constexpr int NKK = K * (K + 1) / 2, NMN = M * (M + 1) / 2;
  for (int idx = threadIdx.x; idx < NMN * NKK; idx += blockDim.x) {
    int mn = idx / NKK, kk = idx % NKK, m = 0, n = 0, k = 0, k2 = 0;
    for (int t = mn; t >= M - m; t -= M - m, m++) {}
    { int t = mn; for (int j = 0; j < m; j++) t -= M - j; n = m + t; }
    for (int t = kk; t >= K - k; t -= K - k, k++) {}
    { int t = kk; for (int j = 0; j < k; j++) t -= K - j; k2 = k + t; }
    const int f1 = k * (M + 1) + m + 1, f2 = k2 * (M + 1) + n + 1, f3 = k2 * (M + 1) + m + 1, f4 = k * (M + 1) + n + 1;
    double q1 = 0, q2 = 0;
    for (int p = 0; p < a.P4; p++) {
      const double* gp = g + p * a.QS;
      q1 = fma(gp[f1], gp[f2], q1);
      q2 = fma(gp[f3], gp[f4], q2);
    }
    Qs[idx] = (k == k2) ? q1 : q1 + q2;
  }
  __syncthreads();
}
// upper triangle of G_i from the forms and z
template <int K, int M>
__device__ __forceinline__ void gram_from_forms(const double* Qs, const double (&z)[K], double (&G)[M][M]) {
  constexpr int NKK = K * (K + 1) / 2;
  double zz[NKK];
  {
    int t = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int k2 = k; k2 < K; k2++, t++) zz[t] = z[k] * z[k2];
  }
  int mn = 0;
#pragma unroll
  for (int m = 0; m < M; m++)
#pragma unroll
    for (int q = m; q < M; q++, mn++) {
      double t = 0;
#pragma unroll
      for (int kk = 0; kk < NKK; kk++) t = fma(Qs[mn * NKK + kk], zz[kk], t);
      G[m][q] = t;
    }
}

template <int K, int M, bool COV, int V, bool RG, bool CPO = false>
#ifndef BF_CHI_MINB
#define BF_CHI_MINB 4      // resident blocks per SM targeted by the V = 2 chi kernel (3: 68 us, 4: 66 us, 5: 83 us, 6: 134 us -- spills)
#endif
#ifndef BF_CHI1_MINB
#define BF_CHI1_MINB 8
#endif
__global__ void __launch_bounds__(PF_THREADS, (V == 1 && !RG) ? BF_CHI1_MINB : (RG ? 4 : BF_CHI_MINB)) chi_kernel(const PassArgs a) {
  extern __shared__ double g[];
  build_log_table();
  stage_globals(a, g);
  // Common basis without covariates: u_m = sum_k z_k phi~_km, so the per-function Gram is a quadratic form in z,
  //   G_i[m][n] = u_m . u_n = sum_{k <= k'} z_k z_k' Q[mn][kk'],   Q[mn][kk'] = phi~_km . phi~_k'n (+ phi~_k'm . phi~_kn, k != k'),
  // with Q a property of the globals only: the block builds it once, and the row loop keeps just the data-dependent
  // part (the residual d = c~ - mu~ and s[k][m] = phi~_km . d, from which r[m] = sum_k z_k s[k][m]): 4 + K M
  // multiply-adds per row instead of 1 + K + K M + M + M(M+1)/2.
  constexpr bool GQ = !COV && !RG;
  constexpr int NKK = K * (K + 1) / 2, NMN = M * (M + 1) / 2;
  __shared__ __align__(16) double Qs[GQ ? NMN * NKK : 1];
  if constexpr (GQ) build_gram_forms<K, M>(a, g, Qs);
  double red[1] = {0};
  const double bs = a.beta / (a.sigma_dev ? *a.sigma_dev : a.sigma_sq);
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    FnState<K, M, COV, V> st;
    st.load(a, i0);
    RowStream<V> rows;
    if constexpr (!RG) rows.begin(a.Ct + i0, a.ld);
    // the normals do not depend on the state: generated first, while the loads above are in flight
    double eps[V][M];
    if (!CPO && !a.eps) {
#pragma unroll
      for (int v = 0; v < V; v++) {
        // (M+1)/2 Box-Muller pairs, one Philox block each (three words used)
        const uint64_t gi = a.global_offset + (uint64_t)(i0 + v);
#pragma unroll
        for (int pr = 0; pr < (M + 1) / 2; pr++) {
          uint32_t w[4];
          philox_rk(a, gi, RNG_CHI, pr, w);
          double n0, n1;
          fast_box_muller(w[0], w[1], w[2], n0, n1);
          eps[v][2 * pr] = n0;
          if (2 * pr + 1 < M) eps[v][2 * pr + 1] = n1;
        }
      }
    }
    double G[V][M][M], r[V][M], d0[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      d0[v] = 0;
#pragma unroll
      for (int m = 0; m < M; m++) {
        r[v][m] = 0;
#pragma unroll
        for (int q = 0; q < M; q++) G[v][m][q] = 0;
      }
    }
    if constexpr (GQ) {
      double sk[V][K][M];
#pragma unroll
      for (int v = 0; v < V; v++)
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
          for (int m = 0; m < M; m++) sk[v][k][m] = 0;
      constexpr int QSc = (K * (M + 1) + 1) & ~1;
      Coef<K, M, false, V> cf;
      rows.run(a.P4, [&](int p, const double (&c)[V]) {
        cf.load(g + p * QSc, 0, st.x);
#pragma unroll
        for (int v = 0; v < V; v++) {
          double dres = c[v];
#pragma unroll
          for (int k = 0; k < K; k++) dres = fma(-st.z[v][k], cf.get(v, k, 0), dres);
          d0[v] = fma(dres, dres, d0[v]);
#pragma unroll
          for (int k = 0; k < K; k++)
#pragma unroll
            for (int m = 0; m < M; m++) sk[v][k][m] = fma(cf.get(v, k, m + 1), dres, sk[v][k][m]);
        }
      });
#pragma unroll
      for (int v = 0; v < V; v++) {
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
          for (int k = 0; k < K; k++) r[v][m] = fma(st.z[v][k], sk[v][k][m], r[v][m]);
        gram_from_forms<K, M>(Qs, st.z[v], G[v]);
      }
    } else {
    Coef<K, M, COV, V> cf;
    constexpr int NB = RG ? BWMAX : 0;
    BandWin wd[V], wu[V][M];
#pragma unroll
    for (int v = 0; v < V; v++) {
      wd[v].clear();
#pragma unroll
      for (int m = 0; m < M; m++) wu[v][m].clear();
    }
    auto body = [&](int p, const double (&c)[V], const double (&gb)[NB > 0 ? NB : 1][V]) {
      cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
      for (int v = 0; v < V; v++) {
        double dres = c[v], um[M];
#pragma unroll
        for (int m = 0; m < M; m++) um[m] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          dres = fma(-st.z[v][k], cf.get(v, k, 0), dres);
#pragma unroll
          for (int m = 0; m < M; m++) um[m] = fma(st.z[v][k], cf.get(v, k, m + 1), um[m]);
        }
        if constexpr (RG) {       // the same forms through the band of G_i
          double gv[BWMAX];
#pragma unroll
          for (int j = 0; j < BWMAX; j++) gv[j] = gb[j][v];
          d0[v] += band_term_sq(gv, dres, wd[v]);
#pragma unroll
          for (int m = 0; m < M; m++) {
            r[v][m] += band_term(gv, um[m], dres, wu[v][m], wd[v]);
            G[v][m][m] += band_term_sq(gv, um[m], wu[v][m]);
#pragma unroll
            for (int q = m + 1; q < M; q++) G[v][m][q] += band_term(gv, um[m], um[q], wu[v][m], wu[v][q]);
          }
          wd[v].push(dres);
#pragma unroll
          for (int m = 0; m < M; m++) wu[v][m].push(um[m]);
        } else {
          d0[v] = fma(dres, dres, d0[v]);
#pragma unroll
          for (int m = 0; m < M; m++) {
            r[v][m] = fma(um[m], dres, r[v][m]);
#pragma unroll
            for (int q = m; q < M; q++) G[v][m][q] = fma(um[m], um[q], G[v][m][q]);
          }
        }
      }
    };
    if constexpr (RG) {
      stream_rows<V, ROW_CH_RAGGED, NB>(a.Ct, a.Gl, a.bw, a.ld, a.P, i0, body);
    } else {
      const double nogb[1][V] = {};
      rows.run(a.P4, [&](int p, const double (&c)[V]) { body(p, c, nogb); });
    }
    }
    if constexpr (CPO) {
      double rssv[V];
      ldv<V>(a.rss + i0, rssv);
#pragma unroll
      for (int v = 0; v < V; v++) {
        // Cholesky of A = sigma^2 I + G in registers, w = L^{-1} r
        double L[M][M], w[M], logdet = 0, quad = 0;
#pragma unroll
        for (int j = 0; j < M; j++) {
          double dj = G[v][j][j] + a.sigma_sq;
#pragma unroll
          for (int k = 0; k < j; k++) dj = fma(-L[j][k], L[j][k], dj);
          const double lj = sqrt(dj);
          L[j][j] = lj;
          logdet += nl_log(dj);                       // = 2 log L_jj
          double wj = r[v][j];
#pragma unroll
          for (int k = 0; k < j; k++) wj = fma(-L[j][k], w[k], wj);
          w[j] = wj / lj;
          quad = fma(w[j], w[j], quad);
#pragma unroll
          for (int i = j + 1; i < M; i++) {
            double t = G[v][j][i];                      // upper triangle holds G[j][i], j <= i
#pragma unroll
            for (int k = 0; k < j; k++) t = fma(-L[i][k], L[j][k], t);
            L[i][j] = t / lj;
          }
        }
        const double ni = a.ni ? a.ni[i0 + v] : a.npts_common;
        const double lsig = nl_log(a.sigma_sq);
        const double resid = rssv[v] + d0[v];
        const double logl = -0.5 * ni * 1.8378770664093454836 - 0.5 * ((ni - M) * lsig + logdet) - 0.5 * (resid - quad) / a.sigma_sq;
        if (i0 + v < a.n) {
          if (a.logl_out) a.logl_out[i0 + v] = logl;
          if (a.cpo_m) {                                // running log-sum-exp of -logl
            const double x = -logl;
            if (a.cpo_first) { a.cpo_m[i0 + v] = x; a.cpo_s[i0 + v] = 1.0; }
            else {
              const double mo = a.cpo_m[i0 + v], so = a.cpo_s[i0 + v];
              const double mn = fmax(mo, x);
              a.cpo_m[i0 + v] = mn;
              a.cpo_s[i0 + v] = so * exp(mo - mn) + exp(x - mn);
            }
          }
        }
      }
      continue;
    }
    if (a.eps) {
      double t[V];
#pragma unroll
      for (int m = 0; m < M; m++) {
        ldv_cs<V>(a.eps + (size_t)m * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) eps[v][m] = t[v];
      }
    } else {
      if (a.draws_out) {
        double t[V];
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
          for (int v = 0; v < V; v++) t[v] = eps[v][m];
          stv<V>(a.draws_out + (size_t)m * a.ld + i0, t);
        }
      }
    }
    double rssv[V];
    ldv<V>(a.rss + i0, rssv);
#pragma unroll
    for (int v = 0; v < V; v++) {
#pragma unroll
      for (int m = 0; m < M; m++) {
        double w = r[v][m];
#pragma unroll
        for (int q = 0; q < M; q++)
          if (q != m) w = fma(-(q < m ? G[v][q][m] : G[v][m][q]), st.chi[v][q], w);
        const double W = fast_rcp(fma(G[v][m][m], bs, 1.0));        // 1 / (1 + beta G_mm / sigma^2) in (0, 1]
        st.chi[v][m] = fma(W * bs, w, fast_sqrt(W) * eps[v][m]);
      }
      // residual sum of squares with the new chi: rss + |d|^2 - 2 chi'r + chi' G chi
      double quad = 0, lin = 0;
#pragma unroll
      for (int m = 0; m < M; m++) {
        lin = fma(st.chi[v][m], r[v][m], lin);
#pragma unroll
        for (int q = 0; q < M; q++)
          quad = fma(st.chi[v][m] * st.chi[v][q], (q < m ? G[v][q][m] : G[v][m][q]), quad);
      }
      if (i0 + v < a.n) red[0] += rssv[v] + (d0[v] - 2 * lin + quad);
    }
    {
      double t[V];
#pragma unroll
      for (int m = 0; m < M; m++) {
#pragma unroll
        for (int v = 0; v < V; v++) t[v] = st.chi[v][m];
        stv<V>(a.chi + (size_t)m * a.ld + i0, t);
      }
    }
  }
  if constexpr (!CPO) grid_reduce<1>(red, a);
}

// ================================================================= residual sum of squares
// the data pass of updateSigma / calcLikelihood (UpdateSigma.h:36-50, CalculateLikelihood.h:28-42)
template <int K, int M, bool COV, int V, bool RG>
__global__ void __launch_bounds__(PF_THREADS) ssr_kernel(const PassArgs a) {
  extern __shared__ double g[];
  stage_globals(a, g);
  double red[1] = {0};
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    FnState<K, M, COV, V> st;
    st.load(a, i0);
    double acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) acc[v] = 0;
    Coef<K, M, COV, V> cf;
    constexpr int NB = RG ? BWMAX : 0;
    BandWin wr[V];
#pragma unroll
    for (int v = 0; v < V; v++) wr[v].clear();
    auto body = [&](int p, const double (&c)[V], const double (&gb)[NB > 0 ? NB : 1][V]) {
      cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
      for (int v = 0; v < V; v++) {
        double res = c[v];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double at = cf.get(v, k, 0);
#pragma unroll
          for (int m = 0; m < M; m++) at = fma(st.chi[v][m], cf.get(v, k, m + 1), at);
          res = fma(-st.z[v][k], at, res);
        }
        if constexpr (RG) {
          double gv[BWMAX];
#pragma unroll
          for (int j = 0; j < BWMAX; j++) gv[j] = gb[j][v];
          acc[v] += band_term_sq(gv, res, wr[v]);
          wr[v].push(res);
        } else {
          acc[v] = fma(res, res, acc[v]);
        }
      }
    };
    if constexpr (RG) {
      stream_rows<V, ROW_CH_RAGGED, NB>(a.Ct, a.Gl, a.bw, a.ld, a.P, i0, body);
    } else {
      const double nogb[1][V] = {};
      RowStream<V> rows;
      rows.begin(a.Ct + i0, a.ld);
      rows.run(a.P4, [&](int p, const double (&c)[V]) { body(p, c, nogb); });
    }
    double rssv[V];
    ldv<V>(a.rss + i0, rssv);
#pragma unroll
    for (int v = 0; v < V; v++)
      if (i0 + v < a.n) red[0] += rssv[v] + acc[v];
  }
  grid_reduce<1>(red, a);
}

// ------------------------------------------------------------------ dispatch over (K, M, COV)
#ifdef BF_KM_SMALL   /* fast development build: a handful of shapes */
#define BF_KM_CASES(X) X(2, 2) X(2, 3) X(3, 2) X(3, 3) X(4, 3) X(3, 4)
#else
#define BF_KM_CASES(X)                                                                     \
  X(2, 1) X(2, 2) X(2, 3) X(2, 4) X(2, 5) X(2, 6) X(3, 1) X(3, 2) X(3, 3) X(3, 4) X(3, 5) X(3, 6) \
  X(4, 1) X(4, 2) X(4, 3) X(4, 4) X(4, 5) X(4, 6) X(5, 1) X(5, 2) X(5, 3) X(5, 4) X(5, 5) X(5, 6) \
  X(6, 1) X(6, 2) X(6, 3) X(6, 4) X(6, 5) X(6, 6)
#endif

template <int V, typename Kern, typename... Extra>
inline int launch_pass(Kern kern, const PassArgs& a, cudaStream_t s, size_t extra_smem = 0, Extra... extra) {
  size_t smem = (size_t)a.P4 * a.QS * sizeof(double) + extra_smem;
  // resident blocks per SM of this instantiation (queried once), grid = one full wave
  // (kernel, device) -> (smem, blocks per SM): the attribute and the occupancy are per device; several host threads
  // may drive several engines (one per GPU), so the cache is locked
  static std::map<std::pair<const void*, int>, std::pair<size_t, int>> cache;
  static std::mutex cache_mu;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lock(cache_mu);
    const auto key = std::make_pair((const void*)kern, dev);
    auto it = cache.find(key);
    if (it == cache.end() || it->second.first != smem) {
      if (smem > 40 * 1024) {     // dynamic + up to ~6 KB of static shared memory (tables, reduction scratch) must stay within the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
      }
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PF_THREADS, smem) != cudaSuccess || nb < 1) nb = 1;
      cache[key] = std::make_pair(smem, nb);
      it = cache.find(key);
    }
    per_sm = it->second.second;
  }
  int need = pass_grid(a.ld, V);
  int grid = a.sm_count * per_sm - a.grid_reserve;
  if (grid < 1) grid = 1;
  if (grid > need) grid = need;
  if (grid > a.max_blocks) grid = a.max_blocks;
  kern<<<grid, PF_THREADS, smem, s>>>(a, extra...);
  g_launch_count++;
  return (int)cudaGetLastError();
}

#define BF_DISPATCH(KERNEL)                                                        \
  const bool cov = a.D > 0;                                                        \
  switch (K * 16 + M) {                                                            \
    BF_KM_CASES(BF_CASE_##KERNEL)                                                  \
    default: return -2;                                                            \
  }

}  // namespace bf
