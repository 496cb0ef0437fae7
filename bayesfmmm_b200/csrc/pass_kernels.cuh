// pass_kernels.cuh -- the three per-function passes (Z Metropolis step, chi sweep, residual
// sum of squares) over the projected coefficient cache.  One thread owns V (1 or 2) adjacent
// functions; every global access is a 16-byte load/store on coefficient-major (SoA) arrays, so a
// warp reads 512 contiguous bytes per row.  The global coefficients live in shared memory
// (broadcast reads).  K and M are compile-time so all per-function state stays in registers.
#pragma once
#include <unordered_map>
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
// log Z of the current state is read from the cache a.lZ (and written back with Z), every other
// transcendental goes through fastmath.cuh.

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

// generic proposal sampler (rejected fast-path candidate, or K > 4): out of line, rarely executed
template <int K>
__device__ __noinline__ void z_slow_proposal(uint64_t key, uint64_t gi, uint64_t iteration, const double* sh, double* zp,
                                             double* uacc, bool live) {
  RngStream rs(key, gi, iteration, RNG_Z_PROPOSAL_SLOW);
#pragma unroll
  for (int k = 0; k < K; k++) zp[k] = live ? rs.gamma(sh[k]) : 1.0;
  if (K > 4) *uacc = rs.uniform();
}
static __device__ __noinline__ double nl_pow_u(double u, double inv_shape) { return exp(log(u) * inv_shape); }

// log Gamma(tot) for tot = a * sum_k z_k, which equals a up to rounding when the row of Z sums to one:
// second-order Taylor series around a (error |dt|^3 / (6 a^2)), the general routine otherwise
__device__ __forceinline__ double lgamma_near_a(double tot, const PassArgs& a) {
  const double dt = tot - a.a_Z_PM;
  if (fabs(dt) <= 1e-8 * a.a_Z_PM) return fma(dt, fma(0.5 * dt, a.trigam_a, a.digam_a), a.lgam_a);
  return lgamma_shift16(tot, fast_log_nl(tot));     // rows of Z that do not sum to one (rare): the shift is valid for any tot > 0
}

#ifndef BF_Z_MINB
#define BF_Z_MINB 5       // resident blocks per SM the register allocation of the V = 1 common-grid Z kernel targets
                          // (96 registers; measured 8: 122 us, 7: 119, 6: 116, 5: 109, 4: 114 -- spills cost more than warps)
#endif
template <int K, int M, bool COV, int V, bool RG>
__global__ void __launch_bounds__(PF_THREADS, (V == 1 && !RG) ? BF_Z_MINB : 4) z_kernel(const PassArgs a) {
  extern __shared__ double g[];
  build_log_table();
  stage_globals(a, g);
  double red[K + 1];
#pragma unroll
  for (int j = 0; j <= K; j++) red[j] = 0;
  const double hb = a.beta / (2 * a.sigma_sq);
  // persistent blocks: every thread walks the functions with a grid stride, one reduction per block
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    FnState<K, M, COV, V> st;
    st.load(a, i0);
    RowStream<V> rows;
    if constexpr (!RG) rows.begin(a.Ct + i0, a.ld);      // first rows in flight during the proposal
    double zp[V][K], uacc[V], lzo[V][K], lzn[V][K];
    {
      double t[V];
#pragma unroll
      for (int k = 0; k < K; k++) {
        ldv<V>(a.lZ + (size_t)k * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) lzo[v][k] = t[v];
      }
    }
    // ---- proposal
    if (a.gam) {
      double t[V];
#pragma unroll
      for (int k = 0; k < K; k++) {
        ldv_cs<V>(a.gam + (size_t)k * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) zp[v][k] = t[v];
      }
      ldv_cs<V>(a.u + i0, uacc);
#pragma unroll
      for (int v = 0; v < V; v++) {
        double sum = 0;
#pragma unroll
        for (int k = 0; k < K; k++) sum += zp[v][k];
#pragma unroll
        for (int k = 0; k < K; k++) { zp[v][k] = zp[v][k] / sum; lzn[v][k] = fast_log(zp[v][k]); }
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; v++) {
        const uint64_t gi = a.global_offset + (uint64_t)(i0 + v);
        const bool live = (i0 + v) < a.n;
        bool fast = (K <= 4);
        // The random words and the normals do not depend on the state: they are generated first, so the
        // loads of Z, chi and log Z issued above are in flight for a few hundred instructions before their
        // first use (they were 11 % of the kernel's stall samples when the shapes were computed first).
        // Straight-line path: three Philox blocks give two Box-Muller pairs (4 normals), four 32-bit
        // accept uniforms and the Metropolis uniform; one Marsaglia-Tsang candidate per coordinate.
        double nrm[4] = {0, 0, 0, 0}, ua[4] = {0.5, 0.5, 0.5, 0.5};
        if constexpr (K <= 4) {
          uint32_t w0[4], w1[4], w2[4];
          philox_words(a.key, gi, a.iteration, RNG_Z_PROPOSAL, 0, w0);
          philox_words(a.key, gi, a.iteration, RNG_Z_PROPOSAL, 1, w1);
          philox_words(a.key, gi, a.iteration, RNG_Z_PROPOSAL, 2, w2);
          fast_box_muller(w0[0], w0[1], w0[2], nrm[0], nrm[1]);
          if constexpr (K > 2) fast_box_muller(w1[0], w1[1], w1[2], nrm[2], nrm[3]);
          ua[0] = u32(w0[3]); ua[1] = u32(w1[3]); ua[2] = u32(w2[2]); ua[3] = u32(w2[3]);
          uacc[v] = u52(w2[0], w2[1]);
        }
        bool small = false;
        double sh[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
          sh[k] = a.a_Z_PM * st.z[v][k];
          if (!(sh[k] > 0)) sh[k] = 10;                                   // Distributions.h:24-28
          small = small || (sh[k] < 1.0);
        }
        if constexpr (K <= 4) {
          // shapes below 1 (tiny memberships): Gamma(s) = Gamma(s + 1) U^(1/s); the extra uniforms are
          // only generated by warps that contain such a function
          double ub[4] = {0.5, 0.5, 0.5, 0.5};
          if (__any_sync(__activemask(), small)) {
            uint32_t w3[4], w4[4];
            philox_words(a.key, gi, a.iteration, RNG_Z_PROPOSAL, 3, w3);
            philox_words(a.key, gi, a.iteration, RNG_Z_PROPOSAL, 4, w4);
            ub[0] = u53(w3[0], w3[1]); ub[1] = u53(w3[2], w3[3]); ub[2] = u53(w4[0], w4[1]); ub[3] = u53(w4[2], w4[3]);
          }
#pragma unroll
          for (int k = 0; k < K; k++) {
            const bool bo = sh[k] < 1.0;
            const bool ok = gamma_candidate_fast(bo ? sh[k] + 1.0 : sh[k], nrm[k], ua[k], zp[v][k]);
            if (bo) zp[v][k] *= nl_pow_u(ub[k], 1.0 / sh[k]);
            fast = fast && ok;
          }
        }
        if (!fast) {          // a rejected candidate or K > 4: generic sampler on its own stream
          double tsh[K], tzp[K], tu = uacc[v];     // only these copies have their address taken
#pragma unroll
          for (int k = 0; k < K; k++) tsh[k] = sh[k];
          z_slow_proposal<K>(a.key, gi, a.iteration, tsh, tzp, &tu, live);
#pragma unroll
          for (int k = 0; k < K; k++) zp[v][k] = tzp[k];
          uacc[v] = tu;
        }
        double sum = 0;
#pragma unroll
        for (int k = 0; k < K; k++) sum += zp[v][k];
        if (a.draws_out) {
#pragma unroll
          for (int k = 0; k < K; k++) a.draws_out[(size_t)k * a.ld + i0 + v] = zp[v][k];
          a.draws_out[(size_t)K * a.ld + i0 + v] = uacc[v];
        }
        const double rs = (sum > 1e-290 && sum < 1e290) ? fast_rcp(sum) : 1.0 / sum;
#pragma unroll
        for (int k = 0; k < K; k++) { zp[v][k] *= rs; lzn[v][k] = fast_log(zp[v][k]); }
      }
    }
    // ---- squared errors of the current and the proposed state
    double so[V], sn[V];
#pragma unroll
    for (int v = 0; v < V; v++) { so[v] = 0; sn[v] = 0; }
    Coef<K, M, COV, V> cf;
    if constexpr (!RG) {
      const int QSc = COV ? a.QS : ((K * (M + 1) + 1) & ~1);      // compile-time without covariates
      rows.run(a.P4, [&](int p, const double (&c)[V]) {
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
      });
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
    double znew[V][K], lznew[V][K];
#pragma unroll
    for (int v = 0; v < V; v++) {
      double lp = 0;                 // lpdf(z*) - lpdf(z)
      bool nonpos = false;
#pragma unroll
      for (int k = 0; k < K; k++) {
        lp = fma(a.alpha3 * a.pi[k] - 1, lzn[v][k] - lzo[v][k], lp);
        nonpos |= (st.z[v][k] <= 0);
      }
      lp = fma(-hb, sn[v] - so[v], lp);
      double q_new = 0, q_old = 0, lB_new = 0, lB_old = 0, tot_new = 0, tot_old = 0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double al_from_old = a.a_Z_PM * st.z[v][k];   // parameters used to propose the new state
        const double al_from_new = a.a_Z_PM * zp[v][k];     // parameters of the reverse move
        q_new = fma(al_from_old - 1, lzn[v][k], q_new);
        q_old = fma(al_from_new - 1, lzo[v][k], q_old);
        // log(a * z) = log a + log z: both logs are already in registers
        lB_new += lgamma_pos(al_from_old, a.log_a_Z_PM + lzo[v][k]); tot_new += al_from_old;
        lB_old += lgamma_pos(al_from_new, a.log_a_Z_PM + lzn[v][k]); tot_old += al_from_new;
      }
      q_new -= (lB_new - lgamma_near_a(tot_new, a));
      q_old -= (lB_old - lgamma_near_a(tot_old, a));
      double acc = lp + (q_old - q_new);
      if (nonpos) acc = 1;                           // UpdateMixedMembership.h:170-174
      const bool live = (i0 + v) < a.n;
      const bool take = live && (fast_log(uacc[v]) < acc);
      if (a.acc_out && live) a.acc_out[i0 + v] = acc;
#pragma unroll
      for (int k = 0; k < K; k++) {
        znew[v][k] = take ? zp[v][k] : st.z[v][k];
        lznew[v][k] = take ? lzn[v][k] : lzo[v][k];
        if (live) red[k] += lznew[v][k];
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
#pragma unroll
        for (int v = 0; v < V; v++) t[v] = lznew[v][k];
        stv<V>(a.lZ + (size_t)k * a.ld + i0, t);
      }
    }
  }
  grid_reduce<K + 1>(red, a);
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
          philox_words(a.key, gi, a.iteration, RNG_CHI, pr, w);
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

template <int V, typename Kern>
inline int launch_pass(Kern kern, const PassArgs& a, cudaStream_t s, size_t extra_smem = 0) {
  size_t smem = (size_t)a.P4 * a.QS * sizeof(double) + extra_smem;
  // resident blocks per SM of this instantiation (queried once), grid = one full wave
  static std::unordered_map<const void*, std::pair<size_t, int>> cache;   // kernel -> (smem, blocks per SM)
  int dev = 0;
  cudaGetDevice(&dev);
  const void* key = (const void*)((const char*)kern + dev);      // the attribute / occupancy are per device
  auto it = cache.find(key);
  if (it == cache.end() || it->second.first != smem) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PF_THREADS, smem) != cudaSuccess || nb < 1) nb = 1;
    cache[key] = std::make_pair(smem, nb);
    it = cache.find(key);
  }
  const int per_sm = it->second.second;
  int need = pass_grid(a.ld, V);
  int grid = a.sm_count * per_sm;
  if (grid > need) grid = need;
  if (grid > a.max_blocks) grid = a.max_blocks;
  kern<<<grid, PF_THREADS, smem, s>>>(a);
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
