// pass_kernels.cuh -- the three per-function passes (Z Metropolis step, chi sweep, residual
// sum of squares) over the projected coefficient cache.  One thread owns VEC = 2 adjacent
// functions; every global access is a 16-byte load/store on coefficient-major (SoA) arrays, so a
// warp reads 512 contiguous bytes per row.  The global coefficients live in shared memory
// (broadcast reads).  K and M are compile-time so all per-function state stays in registers.
#pragma once
#include "common.cuh"

namespace bf {

// Effective coefficients of one function at basis column p:
//   a[k][0]   = (nu_k + eta_k x_i)[p],   a[k][m+1] = (phi_km + xi_km x_i)[p]     (whitened)
template <int K, int M, bool COV>
struct Coef {
  static constexpr int NV = COV ? VEC : 1;
  double a[NV][K][M + 1];
  __device__ __forceinline__ void load(const double* __restrict__ gs, int D, const double (&x)[VEC][DMAX]) {
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
          for (int v = 0; v < VEC; v++) a[v][k][m] = base;
#pragma unroll
          for (int d = 0; d < DMAX; d++)
            if (d < D) {
              double g = b[1 + d];
#pragma unroll
              for (int v = 0; v < VEC; v++) a[v][k][m] = fma(x[v][d], g, a[v][k][m]);
            }
        }
    }
  }
  __device__ __forceinline__ double get(int v, int k, int m) const { return a[COV ? v : 0][k][m]; }
};

template <int K, int M, bool COV>
struct FnState {
  double z[VEC][K], chi[VEC][M > 0 ? M : 1], x[VEC][DMAX];
  __device__ __forceinline__ void load(const PassArgs& a, int i0) {
#pragma unroll
    for (int k = 0; k < K; k++) { double2 t = ld2(a.Z + (size_t)k * a.ld + i0); z[0][k] = t.x; z[1][k] = t.y; }
#pragma unroll
    for (int m = 0; m < M; m++) { double2 t = ld2(a.chi + (size_t)m * a.ld + i0); chi[0][m] = t.x; chi[1][m] = t.y; }
#pragma unroll
    for (int d = 0; d < DMAX; d++) { x[0][d] = 0; x[1][d] = 0; }
    if constexpr (COV) {
#pragma unroll
      for (int d = 0; d < DMAX; d++)
        if (d < a.D) { double2 t = ld2(a.X + (size_t)d * a.ld + i0); x[0][d] = t.x; x[1][d] = t.y; }
    }
  }
};

__device__ __forceinline__ void stage_globals(const PassArgs& a, double* g) {
  const int tot = a.P * a.QS;
  for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) g[idx] = a.glob[idx];
  __syncthreads();
}

// ================================================================= Z Metropolis step
// updateZ_PM and twins (reference UpdateMixedMembership.h:131-185; lpdf_z :20-50; tempered :91;
// Z_proposal_density :102-113; rdirichlet Distributions.h:22-45; calc_lB :51-61).
// The squared-error terms are evaluated in the whitened coefficient space, where
// ||y - B theta||^2 = rss_i + ||c~_i - theta~||^2 and rss_i cancels in the ratio.
template <int K, int M, bool COV>
__global__ void __launch_bounds__(PF_THREADS) z_kernel(const PassArgs a) {
  extern __shared__ double g[];
  stage_globals(a, g);
  const int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * VEC;
  double red[K + 1];
#pragma unroll
  for (int j = 0; j <= K; j++) red[j] = 0;
  if (i0 < a.ld) {
    FnState<K, M, COV> st;
    st.load(a, i0);
    double zp[VEC][K], uacc[VEC];
    // ---- proposal
    if (a.gam) {
#pragma unroll
      for (int k = 0; k < K; k++) { double2 t = ld2_stream(a.gam + (size_t)k * a.ld + i0); zp[0][k] = t.x; zp[1][k] = t.y; }
      double2 t = ld2_stream(a.u + i0);
      uacc[0] = t.x; uacc[1] = t.y;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        RngStream rs(a.key, a.global_offset + (uint64_t)(i0 + v), a.iteration, RNG_Z_PROPOSAL);
#pragma unroll
        for (int k = 0; k < K; k++) {
          double sh = a.a_Z_PM * st.z[v][k];
          if (sh <= 0) sh = 10;                    // Distributions.h:24-28
          zp[v][k] = (i0 + v < a.n) ? rs.gamma(sh) : 1.0;
        }
        uacc[v] = rs.uniform();
      }
      if (a.draws_out) {
#pragma unroll
        for (int k = 0; k < K; k++) st2(a.draws_out + (size_t)k * a.ld + i0, zp[0][k], zp[1][k]);
        st2(a.draws_out + (size_t)K * a.ld + i0, uacc[0], uacc[1]);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      double sum = 0;
#pragma unroll
      for (int k = 0; k < K; k++) sum += zp[v][k];
#pragma unroll
      for (int k = 0; k < K; k++) zp[v][k] = zp[v][k] / sum;
    }
    // ---- squared errors of the current and the proposed state
    double so[VEC] = {0, 0}, sn[VEC] = {0, 0};
    Coef<K, M, COV> cf;
#pragma unroll 4
    for (int p = 0; p < a.P; p++) {
      double2 c2 = ld2_stream(a.Ct + (size_t)p * a.ld + i0);
      const double c[VEC] = {c2.x, c2.y};
      cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
      for (int v = 0; v < VEC; v++) {
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
    }
    // ---- acceptance
    double znew[VEC][K];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      double lzo[K], lzn[K];
      double lp_old = 0, lp_new = 0;
      bool nonpos = false;
#pragma unroll
      for (int k = 0; k < K; k++) {
        lzo[k] = log(st.z[v][k]);
        lzn[k] = log(zp[v][k]);
        lp_old += (a.alpha3 * a.pi[k] - 1) * lzo[k];
        lp_new += (a.alpha3 * a.pi[k] - 1) * lzn[k];
        nonpos |= (st.z[v][k] <= 0);
      }
      lp_old -= a.beta * (so[v] / (2 * a.sigma_sq));
      lp_new -= a.beta * (sn[v] / (2 * a.sigma_sq));
      double q_new = 0, q_old = 0, lB_new = 0, lB_old = 0, tot_new = 0, tot_old = 0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        double al_from_old = a.a_Z_PM * st.z[v][k];   // parameters used to propose the new state
        double al_from_new = a.a_Z_PM * zp[v][k];     // parameters of the reverse move
        q_new += (al_from_old - 1) * lzn[k];
        q_old += (al_from_new - 1) * lzo[k];
        lB_new += lgamma(al_from_old); tot_new += al_from_old;
        lB_old += lgamma(al_from_new); tot_old += al_from_new;
      }
      q_new -= (lB_new - lgamma(tot_new));
      q_old -= (lB_old - lgamma(tot_old));
      double acc = lp_new - lp_old + q_old - q_new;
      if (nonpos) acc = 1;                           // UpdateMixedMembership.h:170-174
      const bool live = (i0 + v) < a.n;
      const bool take = live && (log(uacc[v]) < acc);
      if (a.acc_out && live) a.acc_out[i0 + v] = acc;
#pragma unroll
      for (int k = 0; k < K; k++) {
        znew[v][k] = take ? zp[v][k] : st.z[v][k];
        if (live) red[k] += take ? lzn[k] : lzo[k];
      }
      if (take) red[K] += 1.0;
    }
#pragma unroll
    for (int k = 0; k < K; k++) st2(a.Z + (size_t)k * a.ld + i0, znew[0][k], znew[1][k]);
  }
  grid_reduce<K + 1>(red, a);
}

// ================================================================= chi sweep (+ post-update SSR)
// updateChi and twins (reference UpdateChi.h:19-64; tempered :116-119).  Per function the M x M
// Gram G[m][n] = ph_m . ph_n and r[m] = ph_m . (y - B mu) are accumulated once (in coefficient
// space), then the reference's sequential m = 0..M-1 sweep is run on them, so chi(i,n) for n < m
// is the already-updated value exactly as in UpdateChi.h:48.
template <int K, int M, bool COV>
__global__ void __launch_bounds__(PF_THREADS) chi_kernel(const PassArgs a) {
  extern __shared__ double g[];
  stage_globals(a, g);
  const int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * VEC;
  double red[1] = {0};
  if (i0 < a.ld) {
    FnState<K, M, COV> st;
    st.load(a, i0);
    double G[VEC][M][M], r[VEC][M], d0[VEC] = {0, 0};
#pragma unroll
    for (int v = 0; v < VEC; v++)
#pragma unroll
      for (int m = 0; m < M; m++) {
        r[v][m] = 0;
#pragma unroll
        for (int q = 0; q < M; q++) G[v][m][q] = 0;
      }
    Coef<K, M, COV> cf;
#pragma unroll 4
    for (int p = 0; p < a.P; p++) {
      double2 c2 = ld2_stream(a.Ct + (size_t)p * a.ld + i0);
      const double c[VEC] = {c2.x, c2.y};
      cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        double dres = c[v], um[M];
#pragma unroll
        for (int m = 0; m < M; m++) um[m] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          dres = fma(-st.z[v][k], cf.get(v, k, 0), dres);
#pragma unroll
          for (int m = 0; m < M; m++) um[m] = fma(st.z[v][k], cf.get(v, k, m + 1), um[m]);
        }
        d0[v] = fma(dres, dres, d0[v]);
#pragma unroll
        for (int m = 0; m < M; m++) {
          r[v][m] = fma(um[m], dres, r[v][m]);
#pragma unroll
          for (int q = m; q < M; q++) G[v][m][q] = fma(um[m], um[q], G[v][m][q]);
        }
      }
    }
    double eps[VEC][M];
    if (a.eps) {
#pragma unroll
      for (int m = 0; m < M; m++) { double2 t = ld2_stream(a.eps + (size_t)m * a.ld + i0); eps[0][m] = t.x; eps[1][m] = t.y; }
    } else {
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        RngStream rs(a.key, a.global_offset + (uint64_t)(i0 + v), a.iteration, RNG_CHI);
#pragma unroll
        for (int m = 0; m < M; m++) eps[v][m] = rs.normal();
      }
      if (a.draws_out) {
#pragma unroll
        for (int m = 0; m < M; m++) st2(a.draws_out + (size_t)m * a.ld + i0, eps[0][m], eps[1][m]);
      }
    }
    const double2 rs2 = ld2(a.rss + i0);
    const double rssv[VEC] = {rs2.x, rs2.y};
#pragma unroll
    for (int v = 0; v < VEC; v++) {
#pragma unroll
      for (int m = 0; m < M; m++) {
        double w = r[v][m];
#pragma unroll
        for (int q = 0; q < M; q++)
          if (q != m) w = fma(-(q < m ? G[v][q][m] : G[v][m][q]), st.chi[v][q], w);
        w = (w * a.beta) / a.sigma_sq;
        double W = 1 + ((G[v][m][m] * a.beta) / a.sigma_sq);
        W = 1 / W;
        st.chi[v][m] = W * w + sqrt(W) * eps[v][m];
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
#pragma unroll
    for (int m = 0; m < M; m++) st2(a.chi + (size_t)m * a.ld + i0, st.chi[0][m], st.chi[1][m]);
  }
  grid_reduce<1>(red, a);
}

// ================================================================= residual sum of squares
// the data pass of updateSigma / calcLikelihood (UpdateSigma.h:36-50, CalculateLikelihood.h:28-42)
template <int K, int M, bool COV>
__global__ void __launch_bounds__(PF_THREADS) ssr_kernel(const PassArgs a) {
  extern __shared__ double g[];
  stage_globals(a, g);
  const int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * VEC;
  double red[1] = {0};
  if (i0 < a.ld) {
    FnState<K, M, COV> st;
    st.load(a, i0);
    double acc[VEC] = {0, 0};
    Coef<K, M, COV> cf;
#pragma unroll 4
    for (int p = 0; p < a.P; p++) {
      double2 c2 = ld2_stream(a.Ct + (size_t)p * a.ld + i0);
      const double c[VEC] = {c2.x, c2.y};
      cf.load(g + p * a.QS, a.D, st.x);
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        double res = c[v];
#pragma unroll
        for (int k = 0; k < K; k++) {
          double at = cf.get(v, k, 0);
#pragma unroll
          for (int m = 0; m < M; m++) at = fma(st.chi[v][m], cf.get(v, k, m + 1), at);
          res = fma(-st.z[v][k], at, res);
        }
        acc[v] = fma(res, res, acc[v]);
      }
    }
    const double2 rs2 = ld2(a.rss + i0);
    if (i0 < a.n) red[0] += rs2.x + acc[0];
    if (i0 + 1 < a.n) red[0] += rs2.y + acc[1];
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

template <typename Kern>
inline int launch_pass(Kern kern, const PassArgs& a, cudaStream_t s) {
  size_t smem = (size_t)a.P * a.QS * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<pass_grid(a.ld), PF_THREADS, smem, s>>>(a);
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
