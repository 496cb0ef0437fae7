// moments_kernels.cu -- updateSigma's and updateChi's data passes as ONE pass over the coefficient cache
// (common basis, no covariates).
//
// updateSigma (UpdateSigma.h:36-50) needs sum_i |y_i - B theta_i|^2 at the current chi; updateChi (UpdateChi.h:19-64)
// needs, per function, r[m] = u_m . (c~ - mu~) and the Gram u_m . u_n with u_m = sum_k z_k phi~_km, mu~ = sum_k z_k nu~_k.
// Both are functions of the same per-function moments
//     d0 = |c~ - mu~|^2,   r[m],   G[m][n] = quadratic form of z with block-constant coefficients (build_gram_forms):
//     |c~ - theta~|^2 = d0 - 2 chi'r + chi'G chi.
// moments_kernel streams the cache once, reduces the residual sum of squares for the sigma^2 draw and leaves
// (r[0..M-1], rss_i + d0) per function in `mom`; chi_draw_kernel then runs the reference's sequential m = 0..M-1 sweep
// from those M + 1 numbers, Z and chi, without touching the cache again.  The moments do not depend on chi or sigma^2,
// so they stay valid between the two calls (the engine invalidates them whenever Z, the globals or the data change and
// falls back to chi_kernel's own pass).
#include <cstdlib>
#include <cstring>

#include "pass_kernels.cuh"

namespace bf {

#ifndef BF_MOM_MINB
#define BF_MOM_MINB 4
#endif
// One thread, at the very end of the pass (out of line: the kernel's register budget is the streaming loop's).
static __device__ __noinline__ void sigma_tail(double* ssr_slot, const SigmaTail* tp) {
  const SigmaTail& t = *tp;
  double ssr = *ssr_slot;
  if (t.world > 1) {                       // one-slot exchange over the peers' mailboxes (protocol of p2p_hook.cu)
    const int par = (int)(t.xseq & 1ull);
    for (int r = 0; r < t.world; r++)
      (reinterpret_cast<double*>(t.peers.box[r] + P2P_HDR) + ((size_t)par * t.world + t.rank) * t.cap)[0] = ssr;
    __threadfence_system();
    for (int r = 0; r < t.world; r++)
      *(reinterpret_cast<volatile unsigned long long*>(t.peers.box[r]) + t.rank) = t.xseq;
    for (int r = 0; r < t.world; r++) {
      volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(t.peers.box[t.rank]) + r;
      unsigned long long spins = 0;
      while (*mine < t.xseq) {
        if (++spins > (1ull << 31)) __trap();
      }
    }
    __threadfence_system();
    const double* src = reinterpret_cast<const double*>(t.peers.box[t.rank] + P2P_HDR) + (size_t)par * t.world * t.cap;
    ssr = 0;
    for (int r = 0; r < t.world; r++) ssr += __ldcg(src + (size_t)r * t.cap);     // rank order: same bits on every rank
    *ssr_slot = ssr;
  }
  RngStream rs(t.key, 0xB200ull, t.iteration, t.purpose);      // the stream and sampler of sigma_draw_kernel / the host draw
  const double b1 = t.scale_ssr * ssr + t.beta0;
  const double r = (1 / b1) * rs.gamma(t.shape);
  const double sig = 1 / r;
  *t.sigma_dev = sig;
  volatile double* host = t.host;
  host[0] = ssr; host[1] = sig;
  __threadfence_system();
  host[2] = t.seq;
}

template <int K, int M, int V>
__global__ void __launch_bounds__(PF_THREADS, BF_MOM_MINB) moments_kernel(const PassArgs a, double* __restrict__ mom,
                                                                          const __grid_constant__ SigmaTail tail) {
  extern __shared__ double g[];
  stage_globals(a, g);
  constexpr int NKK = K * (K + 1) / 2, NMN = M * (M + 1) / 2;
  __shared__ __align__(16) double Qs[NMN * NKK];
  build_gram_forms<K, M>(a, g, Qs);
  double red[1] = {0};
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    // z only: chi is not needed before the row loop is over (it would be live across it)
    double z[V][K];
    {
      double t[V];
#pragma unroll
      for (int k = 0; k < K; k++) {
        ldv<V>(a.Z + (size_t)k * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) z[v][k] = t[v];
      }
    }
    const double x0[V][DMAX] = {};
    RowStream<V> rows;
    rows.begin(a.Ct + i0, a.ld);
    double d0[V], sk[V][K][M];
#pragma unroll
    for (int v = 0; v < V; v++) {
      d0[v] = 0;
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int m = 0; m < M; m++) sk[v][k][m] = 0;
    }
    constexpr int QSc = (K * (M + 1) + 1) & ~1;
    Coef<K, M, false, V> cf;
    rows.run(a.P4, [&](int p, const double (&c)[V]) {
      cf.load(g + p * QSc, 0, x0);
#pragma unroll
      for (int v = 0; v < V; v++) {
        double dres = c[v];
#pragma unroll
        for (int k = 0; k < K; k++) dres = fma(-z[v][k], cf.get(v, k, 0), dres);
        d0[v] = fma(dres, dres, d0[v]);
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
          for (int m = 0; m < M; m++) sk[v][k][m] = fma(cf.get(v, k, m + 1), dres, sk[v][k][m]);
      }
    });
    double rssv[V], rr[M][V], base[V], chi[V][M];
    ldv<V>(a.rss + i0, rssv);
    {
      double t[V];
#pragma unroll
      for (int m = 0; m < M; m++) {
        ldv<V>(a.chi + (size_t)m * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) chi[v][m] = t[v];
      }
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
      double G[M][M], r[M];
#pragma unroll
      for (int m = 0; m < M; m++) {
        r[m] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) r[m] = fma(z[v][k], sk[v][k][m], r[m]);
        rr[m][v] = r[m];
      }
      gram_from_forms<K, M>(Qs, z[v], G);
      double quad = 0, lin = 0;
#pragma unroll
      for (int m = 0; m < M; m++) {
        lin = fma(chi[v][m], r[m], lin);
#pragma unroll
        for (int q = 0; q < M; q++) quad = fma(chi[v][m] * chi[v][q], (q < m ? G[q][m] : G[m][q]), quad);
      }
      base[v] = rssv[v] + d0[v];
      if (i0 + v < a.n) red[0] += rssv[v] + (d0[v] - 2 * lin + quad);
    }
#pragma unroll
    for (int m = 0; m < M; m++) stv<V>(mom + (size_t)m * a.ld + i0, rr[m]);
    stv<V>(mom + (size_t)M * a.ld + i0, base);
  }
  const bool last = grid_reduce_last<1>(red, a);
  if (tail.on && last && threadIdx.x == 0) sigma_tail(a.out, &tail);
}

template <int K, int M, int V>
__global__ void __launch_bounds__(PF_THREADS, 8) chi_draw_kernel(const PassArgs a, const double* __restrict__ mom) {
  extern __shared__ double g[];
  build_log_table();
  stage_globals(a, g);
  constexpr int NKK = K * (K + 1) / 2, NMN = M * (M + 1) / 2;
  __shared__ __align__(16) double Qs[NMN * NKK];
  build_gram_forms<K, M>(a, g, Qs);
  double red[1] = {0};
  const double bs = a.beta / (a.sigma_dev ? *a.sigma_dev : a.sigma_sq);
  for (int i0 = (blockIdx.x * PF_THREADS + threadIdx.x) * V; i0 < a.ld; i0 += gridDim.x * PF_THREADS * V) {
    FnState<K, M, false, V> st;
    st.load(a, i0);
    double rr[M][V], base[V];
#pragma unroll
    for (int m = 0; m < M; m++) ldv_cs<V>(mom + (size_t)m * a.ld + i0, rr[m]);
    ldv_cs<V>(mom + (size_t)M * a.ld + i0, base);
    double eps[V][M];
    if (a.eps) {
      double t[V];
#pragma unroll
      for (int m = 0; m < M; m++) {
        ldv_cs<V>(a.eps + (size_t)m * a.ld + i0, t);
#pragma unroll
        for (int v = 0; v < V; v++) eps[v][m] = t[v];
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; v++) {
        const uint64_t gi = a.global_offset + (uint64_t)(i0 + v);      // same words as chi_kernel
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
#pragma unroll
    for (int v = 0; v < V; v++) {
      double G[M][M];
      gram_from_forms<K, M>(Qs, st.z[v], G);
#pragma unroll
      for (int m = 0; m < M; m++) {
        double w = rr[m][v];
#pragma unroll
        for (int q = 0; q < M; q++)
          if (q != m) w = fma(-(q < m ? G[q][m] : G[m][q]), st.chi[v][q], w);
        const double W = fast_rcp(fma(G[m][m], bs, 1.0));        // 1 / (1 + beta G_mm / sigma^2) in (0, 1]
        st.chi[v][m] = fma(W * bs, w, fast_sqrt(W) * eps[v][m]);
      }
      double quad = 0, lin = 0;
#pragma unroll
      for (int m = 0; m < M; m++) {
        lin = fma(st.chi[v][m], rr[m][v], lin);
#pragma unroll
        for (int q = 0; q < M; q++) quad = fma(st.chi[v][m] * st.chi[v][q], (q < m ? G[q][m] : G[m][q]), quad);
      }
      if (i0 + v < a.n) red[0] += base[v] + (quad - 2 * lin);
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
  grid_reduce<1>(red, a);
}

constexpr int MV = 2, DV = 1;
#define BF_CASE_moments(KK, MM) \
  case KK * 16 + MM: return cov ? -3 : launch_pass<MV>(moments_kernel<KK, MM, MV>, a, s, 0, mom, tl);
#define BF_CASE_chidraw(KK, MM) \
  case KK * 16 + MM: return cov ? -3 : launch_pass<DV>(chi_draw_kernel<KK, MM, DV>, a, s, 0, (const double*)mom);

int launch_moments(const PassArgs& a, int K, int M, double* mom, const SigmaTail* tail, cudaStream_t s) {
  SigmaTail tl;
  if (tail) tl = *tail; else { std::memset(&tl, 0, sizeof(tl)); }
  BF_DISPATCH(moments)
}
int launch_chi_draw(const PassArgs& a, int K, int M, const double* mom, cudaStream_t s) {
  BF_DISPATCH(chidraw)
}
}  // namespace bf
