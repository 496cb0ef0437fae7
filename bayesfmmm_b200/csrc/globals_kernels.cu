// globals_kernels.cu -- the small updates of the sweep on the device, so that a sweep is a queue of
// kernels with no host round trip on its critical path:
//
//   draw_blocks_kernel   updatePhi then updateNu (UpdatePhi.h:23-89, UpdateNu.h:24-74; MV :190-249, :160-204): the
//                        Gaussian block draws from the reduced statistics W'W, C~'W, in the reference's sequential
//                        block order, then the whitened coefficients the pass kernels stage (glob)
//   sigma_kernel         updateSigma's draw behind the SSR pass (UpdateSigma.h:47-53, tempered :98-107)
//   pi_alpha_kernel      updatePi_PM / updateAlpha3 (UpdatePi.h:84-116, UpdateAlpha3.h:36-63) on a side stream, as soon as
//                        sum_i log Z_ik is reduced: they feed the NEXT sweep's Z step
//   priors_kernel        updateDelta, updateA, updateGamma, updateTau (UpdateDelta.h:17-66, UpdateA.h:58-135,
//                        UpdateGamma.h:17-38, UpdateTau.h:18-68) on the side stream: they feed the NEXT sweep's block draws
//
// All three run the statements of globals_core.cuh -- the code the host loop runs -- on the same Philox streams, one
// device thread per independent unit (a block's factorisation, a delta chain, a gamma row, ...).  The side-stream
// kernels use 64 threads x <= 64 registers: that fits into the registers a resident wave of the Z, SSR and statistics
// kernels leaves free on an SM (4096 of 65536), so they run beside a pass kernel without displacing one of its blocks.
#include <cuda_runtime.h>

#include "common.cuh"
#include "globals_core.cuh"
#include <mutex>
#include <set>

#include "globals_dev.h"

namespace bf {

// ------------------------------------------------------------------ Gaussian block draws
// One thread block of DB_THREADS threads, one WARP per Gaussian block for everything that does not depend on the
// coefficients being drawn (the blocks' precisions depend on the statistics and the priors only):
//   A  statistics, coefficients, Gram matrix into shared memory; the standard normals (one thread per coefficient)
//   B  the banded precisions  Prec_a = beta S_aa G / sigma^2 + Prior_a   (one thread per band entry)
//   C  banded reverse Cholesky  Prec_a = U_a U_a'  (a warp per block: the entries of a column on different lanes)
//   D  T_a = beta Prec_a^-1 G / sigma^2 column by column (lane k solves  U U' y = beta G[:, k] / sigma^2),
//      h_a = beta Prec_a^-1 (B'Y'W)_a / sigma^2 + U_a^-T z_a   (U^-T z = chol_lower(Prec^-1) z, the reference's
//      mvnrnd map, UpdateNu.h:69)
// so that the sequential part -- block a needs the coefficients of the blocks drawn before it, the reference's
// Gauss-Seidel order -- is ONE matrix-vector product per block,
//   c_a = h_a - T_a v,   v = sum_{b != a} S_ab c_b = M_a - S_aa c_a,   then  M_b += S_ab (c_a_new - c_a_old)  for all b,
// run by one warp with row i of T_a in the registers of lane i and v passed around by shuffles: no barrier and no
// shared-memory round trip on the chain.  Diagonal precisions (multivariate model) need no matrices at all.
constexpr int DB_THREADS = 384;
constexpr int DB_WARPS = DB_THREADS / 32;
constexpr int DB_PMAX = 32;             // coefficients per block with a banded (non-diagonal) precision
constexpr int DB_PITCH = 33;

__global__ void __launch_bounds__(DB_THREADS) draw_blocks_kernel(const DrawArgs a) {
  extern __shared__ double sm[];
  const GlobalsView& g = a.g;
  const int K = g.K, P = g.P, M = g.M, q = a.q, ldb = a.hbmax + 1;
  const int nphi = a.do_phi ? K * M : 0, nnu = a.do_nu ? K : 0, nb = nphi + nnu;
  const bool diag = a.hbmax == 0;
  double* S = sm;                          // q x q   W'W
  double* R = S + q * q;                   // P x q   B'Y'W (un-whitened), column f at R + f * P
  double* C = R + P * q;                   // q x P   current coefficients, feature f at C + f * P
  double* Mt = C + q * P;                  // q x P   M_b = sum_a S_ab c_a
  double* Zs = Mt + q * P;                 // nb x P  standard normals
  double* H = Zs + (size_t)nb * P;         // nb x P  h_a
  double* Ub = H + (size_t)nb * P;         // nb x (P * ldb) band of Prec_a, then of U_a
  double* rd = Ub + (size_t)nb * P * ldb;  // nb x P  1 / U_a[i][i]
  double* Gs = rd + (size_t)nb * P;        // P x P basis Gram (banded precisions only)
  double* Cv = Gs + (diag ? 0 : P * P);    // nb x P x DB_PITCH  columns of the solves: T_a[i][k] at (a P + i) DB_PITCH + k
                                           // (odd pitch: lanes = columns while solving, lanes = rows in the chain, both conflict-free)
  __shared__ int s_fail;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_fail = 0;
  int nclk = 0;
  auto stamp = [&]() { if (a.clk && tid == 0) a.clk[nclk] = clock64(); nclk++; };
  stamp();
  const double sigma_sq = *a.sigma_dev;
  const double sc = a.beta / sigma_sq;
  const double* WtW = a.stats + K + 3;
  const double* CtW = WtW + q * q;         // Pc x q, whitened
  // ---- A
  for (int i = tid; i < q * q; i += DB_THREADS) S[i] = WtW[i];
  if (!diag) for (int i = tid; i < P * P; i += DB_THREADS) Gs[i] = a.G[i];
  // B'Y'W = L (C~'W): L is P x Pc, lower banded (half bandwidth hbL) when Pc == P; identity model: L = I
  for (int i = tid; i < P * q; i += DB_THREADS) {
    const int f = i / P, r = i % P;
    double s = 0;
    if (g.identity) s = CtW[f * a.Pc + r];
    else {
      const int kmax = (a.Pc == P) ? r + 1 : a.Pc, kmin = (a.Pc == P) ? (r - a.hbL > 0 ? r - a.hbL : 0) : 0;
      for (int k = kmin; k < kmax; k++) s += a.L[k * P + r] * CtW[f * a.Pc + k];
    }
    R[i] = s;
  }
  for (int i = tid; i < q * P; i += DB_THREADS) {
    const int f = i / P, p = i % P, k = f / (M + 1), mm = f % (M + 1);
    C[i] = mm == 0 ? g.nu_(k, p) : g.Phi_(k, p, mm - 1);
  }
  // the normals of every block: element (block, coefficient) of the block type's stream
  for (int i = tid; i < nb * P; i += DB_THREADS) {
    const int t = i / P, p = i % P;
    const bool is_phi = t < nphi;
    auto st = a.rng.open(is_phi ? HP_PHI : HP_NU, ((uint64_t)(is_phi ? t : t - nphi) << 12) + (uint64_t)p);
    Zs[i] = st.normal();
  }
  __syncthreads();
  stamp();
  // ---- B
  for (int i = tid; i < nb * P * ldb; i += DB_THREADS) {
    const int t = i / (P * ldb), r = (i / ldb) % P, d = i % ldb, c = r + d;
    const bool is_phi = t < nphi;
    const int j = is_phi ? t / M : t - nphi, m = is_phi ? t % M : 0;
    const int f = j * (M + 1) + (is_phi ? m + 1 : 0);
    const int hb = is_phi ? a.hbG : (a.hbG > g.hbP ? a.hbG : g.hbP);
    double v = 0;
    if (c < P && d <= hb) {
      const double gg = g.identity ? (d == 0 ? 1.0 : 0.0) : Gs[c * P + r];
      double pr;
      if (is_phi) {
        pr = 0;
        if (d == 0) {
          double tt = 1;                                   // tilde_tau cumprod, BFMMM.h:1254-1259
          for (int mm = 0; mm <= m; mm++) tt *= g.delta_(j, mm);
          pr = tt * g.gamma_(j, r, m);
        }
      } else pr = g.identity ? (d == 0 ? 1 / g.tau[j] : 0.0) : g.tau[j] * g.Pmat[c * P + r];
      v = sc * S[f * q + f] * gg + pr;
    }
    Ub[i] = v;
  }
  for (int i = tid; i < q * P; i += DB_THREADS) {
    const int b = i / P, p = i % P;
    double s = 0;
    for (int f = 0; f < q; f++) s = fma(S[f * q + b], C[f * P + p], s);
    Mt[i] = s;
  }
  __syncthreads();
  stamp();
  if (diag) {
    // diagonal precisions: U = sqrt(Prec), h_a = sc R_a / Prec + z / sqrt(Prec), T_a = sc / Prec (kept in rd as 1 / Prec)
    for (int i = tid; i < nb * P; i += DB_THREADS) {
      const int t = i / P, p = i % P;
      const bool is_phi = t < nphi;
      const int j = is_phi ? t / M : t - nphi, m = is_phi ? t % M : 0;
      const int f = j * (M + 1) + (is_phi ? m + 1 : 0);
      const double prec = Ub[i];
      if (!(prec > 0)) s_fail = 1;
      const double ri = rsqrt(prec);
      rd[i] = ri * ri;
      H[i] = fma(sc * rd[i], R[f * P + p], ri * Zs[i]);
    }
    __syncthreads();
  } else {
    // ---- C, D, E: one warp per block
    for (int t = warp; t < nb; t += DB_WARPS) {
      double* A = Ub + (size_t)t * P * ldb;
      double* r = rd + (size_t)t * P;
      const bool is_phi = t < nphi;
      const int j = is_phi ? t / M : t - nphi, m = is_phi ? t % M : 0;
      const int f = j * (M + 1) + (is_phi ? m + 1 : 0);
      const int hb = a.hbmax;
      // C: column jj of U: lane d holds entry U[jj - d][jj], d = 0..hb (U[i][jj] = (A[i][jj] - sum_{k > jj} U[i][k] U[jj][k]) / U[jj][jj])
      bool bad = false;
      for (int jj = P - 1; jj >= 0; jj--) {
        const int i = jj - lane;                      // row of this lane's entry
        double tv = 0;
        if (lane <= hb && i >= 0) {
          tv = A[i * ldb + lane];
          const int km = (i + hb < P - 1 ? i + hb : P - 1);
          for (int k = jj + 1; k <= km; k++) tv = fma(-A[i * ldb + (k - i)], A[jj * ldb + (k - jj)], tv);
        }
        const double sdiag = __shfl_sync(0xffffffffu, tv, 0);
        bad = bad || !(sdiag > 0);
        const double inv = rsqrt(sdiag);
        if (lane == 0) { A[jj * ldb] = sdiag * inv; r[jj] = inv; }
        else if (lane <= hb && i >= 0) A[i * ldb + lane] = tv * inv;
        __syncwarp();
      }
      if (bad) s_fail = 1;
      // D: lane k < P solves  U U' y = sc G[:, k]  -> column k of T_a = sc Prec_a^-1 G;  lane P solves
      // U U' y = sc (B'Y'W)_a  -> the mean part of h_a;  lane 31 the forward solve  U' x = z  (= chol_lower(Cov_a) z)
      double* Y = Cv + (size_t)t * P * DB_PITCH;      // Y[i * DB_PITCH + lane]
      for (int i = 0; i < P; i++) Y[i * DB_PITCH + lane] = 0.0;      // columns beyond P + 1 stay zero (the chain reads 32)
      __syncwarp();
      if (lane <= P) {
        const int c = lane;
        for (int i = P - 1; i >= 0; i--) {            // U w = b
          double s = sc * (c < P ? Gs[c * P + i] : R[f * P + i]);
          const int k1 = i + hb < P - 1 ? i + hb : P - 1;
#pragma unroll 4
          for (int k = i + 1; k <= k1; k++) s = fma(-A[i * ldb + (k - i)], Y[k * DB_PITCH + c], s);
          Y[i * DB_PITCH + c] = s * r[i];
        }
        for (int i = 0; i < P; i++) {                 // U' y = w, in place
          double s = Y[i * DB_PITCH + c];
          const int k0 = i - hb > 0 ? i - hb : 0;
#pragma unroll 4
          for (int k = k0; k < i; k++) s = fma(-A[k * ldb + (i - k)], Y[k * DB_PITCH + c], s);
          Y[i * DB_PITCH + c] = s * r[i];
        }
      }
      if (lane == 31) {                               // U' x = z
        double* x = H + (size_t)t * P;
        const double* z = Zs + (size_t)t * P;
        for (int i = 0; i < P; i++) {
          double s = z[i];
          const int k0 = i - hb > 0 ? i - hb : 0;
          for (int k = k0; k < i; k++) s = fma(-A[k * ldb + (i - k)], x[k], s);
          x[i] = s * r[i];
        }
      }
      __syncwarp();
      // h_a = mean part (column P of the solves) + U^-T z
      if (lane < P) H[(size_t)t * P + lane] += Y[lane * DB_PITCH + P];
    }
    __syncthreads();
  }
  stamp();
  stamp();
  // ---- the draws, block after block in the reference's order (Phi: j outer, m inner; then nu: j): warp 0
  if (!s_fail && warp == 0) {
    for (int t = 0; t < nb; t++) {
      const bool is_phi = t < nphi;
      const int j = is_phi ? t / M : t - nphi, m = is_phi ? t % M : 0;
      const int f = j * (M + 1) + (is_phi ? m + 1 : 0);
      const double saa = S[f * q + f];
      if (diag) {
        for (int p = lane; p < P; p += 32) {
          const double v = fma(-saa, C[f * P + p], Mt[f * P + p]);
          const double x = fma(-sc * rd[(size_t)t * P + p], v, H[(size_t)t * P + p]);
          const double dl = x - C[f * P + p];
          C[f * P + p] = x;
          if (is_phi) g.Phi_(j, p, m) = x; else g.nu_(j, p) = x;
          for (int b = 0; b < q; b += 4) {
            double sv[4], mv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { sv[u] = b + u < q ? S[f * q + b + u] : 0.0; mv[u] = b + u < q ? Mt[(b + u) * P + p] : 0.0; }
#pragma unroll
            for (int u = 0; u < 4; u++) if (b + u < q) Mt[(b + u) * P + p] = fma(sv[u], dl, mv[u]);
          }
        }
      } else {
        // row `lane` of T_a into registers (rows >= P are never read: their lanes idle), v around by shuffles
        double trow[DB_PMAX];
        const double* Ti = Cv + ((size_t)t * P + (lane < P ? lane : 0)) * DB_PITCH;
#pragma unroll
        for (int k = 0; k < DB_PMAX; k++) trow[k] = Ti[k];
        const double cold = lane < P ? C[f * P + lane] : 0.0;
        const double v = lane < P ? fma(-saa, cold, Mt[f * P + lane]) : 0.0;
        double x4[4] = {lane < P ? H[(size_t)t * P + lane] : 0.0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < DB_PMAX; k++) x4[k & 3] = fma(-trow[k], __shfl_sync(0xffffffffu, v, k), x4[k & 3]);     // v is zero on the lanes beyond P
        const double x = (x4[0] + x4[1]) + (x4[2] + x4[3]);
        if (lane < P) {
          const double dl = x - cold;
          C[f * P + lane] = x;
          if (is_phi) g.Phi_(j, lane, m) = x; else g.nu_(j, lane) = x;
          for (int b = 0; b < q; b += 4) {                 // M_b += S_ab (c_a_new - c_a_old), four rows at a time
            double sv[4], mv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { sv[u] = b + u < q ? S[f * q + b + u] : 0.0; mv[u] = b + u < q ? Mt[(b + u) * P + lane] : 0.0; }
#pragma unroll
            for (int u = 0; u < 4; u++) if (b + u < q) Mt[(b + u) * P + lane] = fma(sv[u], dl, mv[u]);
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  stamp();
  if (tid == 0 && s_fail) *a.err = 1;
  // ---- whitened coefficients for the pass kernels: glob[p][f] = (L' c_f)[p], zero rows / columns of padding
  for (int i = tid; i < a.P4 * a.QS; i += DB_THREADS) {
    const int p = i / a.QS, f = i % a.QS;
    double s = 0;
    if (p < a.Pc && f < q) {
      if (g.identity) s = C[f * P + p];
      else {
        const int r0 = (a.Pc == P) ? p : 0, r1 = (a.Pc == P) ? (p + a.hbL < P - 1 ? p + a.hbL : P - 1) : P - 1;
        for (int r = r0; r <= r1; r++) s += a.L[p * P + r] * C[f * P + r];
      }
    }
    a.glob[i] = s;
  }
  stamp();
}

size_t draw_blocks_smem(const DrawArgs& a) {
  const size_t K = a.g.K, P = a.g.P, M = a.g.M, q = a.q, nb = K * M + K, ldb = a.hbmax + 1;
  size_t n = q * q + 3 * P * q + nb * P * (3 + ldb);
  if (a.hbmax > 0) n += P * P + nb * P * DB_PITCH;
  return sizeof(double) * (n + 2);
}
int launch_draw_blocks(const DrawArgs& a, cudaStream_t s) {
  if (a.hbmax > 0 && a.g.P > DB_PMAX - 2) return 2;
  const size_t smem = draw_blocks_smem(a);
  if (smem > 48 * 1024) {               // once per device (the attribute is per context)
    static std::set<int> configured;
    static std::mutex configured_mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(configured_mu);
    if (!configured.count(dev)) {
      if (cudaFuncSetAttribute(draw_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 1;
      configured.insert(dev);
    }
  }
  draw_blocks_kernel<<<1, DB_THREADS, smem, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ sigma^2, and pi / alpha_3
__global__ void sigma_kernel(const SigmaPiArgs a) {
  auto st = a.rng.open(HP_SIGMA, 0);
  const double ssr = a.stats[a.g.K + 1];
  const double b1 = a.scale_ssr * ssr + a.g.h.beta_0;
  const double r = (1 / b1) * st.gamma(a.shape);
  *a.sigma_dev = 1 / r;
}
int launch_sigma(const SigmaPiArgs& a, cudaStream_t s) {
  sigma_kernel<<<1, 1, 0, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}
// updatePi_PM then updateAlpha3 from sum_i log Z_ik (side stream: the next sweep's Z step needs them)
__global__ void __launch_bounds__(64, 16) pi_alpha_kernel(const SigmaPiArgs a) {
  __shared__ double s_gam[8];
  const int tid = threadIdx.x;
  const GlobalsView& g = a.g;
  if (tid < g.K) s_gam[tid] = core_pi_gamma(g, a.rng, tid);
  __syncthreads();
  if (tid == 0) {
    auto st = a.rng.open(HP_PI, (uint64_t)g.K);
    core_pi_accept(g, a.stats, s_gam, st.uniform());
    core_update_alpha3(g, a.rng, a.stats);
  }
}
int launch_pi_alpha(const SigmaPiArgs& a, cudaStream_t s) {
  pi_alpha_kernel<<<1, 64, 0, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ delta, A, gamma, tau
__global__ void __launch_bounds__(64, 16) priors_kernel(const PriorsArgs a) {
  const int tid = threadIdx.x;
  const GlobalsView& g = a.g;
  if (a.do_phi && tid < g.K) core_update_delta_k(g, a.rng, tid);                    // delta sees the new Phi, the old gamma and A
  if (tid >= 32 && tid < 32 + g.K) core_update_tau_k(g, a.rng, tid - 32);           // tau sees the new nu
  __syncthreads();
  if (a.do_phi) {
    if (tid < 2 * g.K) core_update_A_one(g, a.rng, tid / 2, tid % 2);               // A sees the new delta
    for (int idx = tid; idx < g.K * g.P; idx += 64) core_update_gamma_row(g, a.rng, idx / g.P, idx % g.P);   // gamma: new delta, new Phi
  }
}
int launch_priors(const PriorsArgs& a, cudaStream_t s) {
  priors_kernel<<<1, 64, 0, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace bf
