// host_sampler.cu -- the host loop above the engine ABI (include/bfmmm_sampler.h): the
// reference's driver loops restated without Rcpp/Armadillo.  Pure host C++ (compiled by nvcc
// only so that it shares common.cuh's Philox generator with the device code).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bfmmm_io.h"
#include "../../include/bfmmm_sampler.h"
#include "common.cuh"
#include "engine_internal.h"
#include "globals_core.cuh"
#include "globals_dev.h"

namespace bf_host { bool chol_upper_rev(int n, const double* A, double* U, int hb); }

namespace {

// ------------------------------------------------------------------ host random numbers
// (the purposes of the global Philox streams, HP_*, live in globals_core.cuh)
using namespace bf;   // HP_*, GlobalsView, core_update_*

struct HostRng {
  uint64_t key = 0, iteration = 0;
  std::deque<double> tape;
  bool use_tape = false;
  bool tape_underrun = false;      // an update asked for more injected draws than the tape held: the caller fails
  bf::RngStream stream{0, 0, 0, 0};
  void open(uint32_t purpose) { stream = bf::RngStream(key, 0xB200ull, iteration, purpose); }
  double pop() {
    if (tape.empty()) { tape_underrun = true; return std::numeric_limits<double>::quiet_NaN(); }
    double v = tape.front(); tape.pop_front(); return v;
  }
  double uniform() { return use_tape ? pop() : stream.uniform(); }
  double gamma(double shape) { return use_tape ? pop() : stream.gamma(shape); }   // Gamma(shape, 1)
};
// the interface globals_core.cuh's templates draw through: a stream per (purpose, element), or the tape of injected
// draws (consumed in call order, the element is ignored)
struct HostStreamRef {
  HostRng* r;
  bf::RngStream st;
#ifdef __CUDA_ARCH__
  __host__ __device__ double normal() { return 0; }
  __host__ __device__ double uniform() { return 0; }
  __host__ __device__ double gamma(double) { return 0; }
#else
  __host__ __device__ double normal() { return r->use_tape ? r->pop() : st.normal(); }
  __host__ __device__ double uniform() { return r->use_tape ? r->pop() : st.uniform(); }
  __host__ __device__ double gamma(double shape) { return r->use_tape ? r->pop() : st.gamma(shape); }
#endif
};
struct HostRngAdapter {
  HostRng* r;
  __host__ __device__ HostStreamRef open(uint32_t purpose, uint64_t element) const {
    return HostStreamRef{r, bf::RngStream(r->key, 0xB200ull + (element << 16), r->iteration, purpose)};
  }
};

// ------------------------------------------------------------------ small dense linear algebra (column-major)
using vecd = std::vector<double>;

// pseudo-inverse of a symmetric matrix (cyclic Jacobi); used when the precision is singular
void pinv_sym_jacobi(int n, const double* A, double* Ainv) {
  vecd a(A, A + (size_t)n * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)q * n + p] * a[(size_t)q * n + p];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[(size_t)q * n + p];
        if (apq == 0.0) continue;
        double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) { double x = a[(size_t)p * n + k], y = a[(size_t)q * n + k]; a[(size_t)p * n + k] = c * x - s * y; a[(size_t)q * n + k] = s * x + c * y; }
        for (int k = 0; k < n; k++) { double x = a[(size_t)k * n + p], y = a[(size_t)k * n + q]; a[(size_t)k * n + p] = c * x - s * y; a[(size_t)k * n + q] = s * x + c * y; }
        for (int k = 0; k < n; k++) { double x = V[(size_t)p * n + k], y = V[(size_t)q * n + k]; V[(size_t)p * n + k] = c * x - s * y; V[(size_t)q * n + k] = s * x + c * y; }
      }
  }
  double lmax = 0;
  for (int i = 0; i < n; i++) lmax = std::max(lmax, std::fabs(a[(size_t)i * n + i]));
  double tol = n * lmax * std::numeric_limits<double>::epsilon();
  std::fill(Ainv, Ainv + (size_t)n * n, 0.0);
  for (int e = 0; e < n; e++) {
    double lam = a[(size_t)e * n + e];
    if (std::fabs(lam) <= tol) continue;
    for (int j = 0; j < n; j++) {
      double vj = V[(size_t)e * n + j] / lam;
      for (int i = 0; i < n; i++) Ainv[(size_t)j * n + i] += V[(size_t)e * n + i] * vj;
    }
  }
}

// ---- contiguous-inner-loop variants used by the per-sweep block draws (P x P, P ~ 20..400) ----
// Cholesky factor of a symmetric matrix, ROW-major lower triangle (row i = L[i*n .. i*n+i])
bool chol_rows(int n, const double* A, double* L) {
  for (int i = 0; i < n; i++) {
    double* li = L + (size_t)i * n;
    for (int j = 0; j <= i; j++) {
      const double* lj = L + (size_t)j * n;
      double s = A[(size_t)j * n + i];
      for (int k = 0; k < j; k++) s -= li[k] * lj[k];
      if (j == i) {
        if (!(s > 0)) return false;
        li[i] = std::sqrt(s);
      } else {
        li[j] = s / lj[j];
      }
    }
    for (int j = i + 1; j < n; j++) li[j] = 0.0;
  }
  return true;
}
// "Reverse" Cholesky A = U U' with U UPPER triangular, row-major (row i = U[i*n + i .. i*n + n-1]).
// Why: the reference draws  x = C b + chol_lower(C) z  with C = A^{-1} (UpdateNu.h:67-69, UpdatePhi.h:79-82).
// With A = U U',  C = U^{-T} U^{-1} = (U^{-T})(U^{-T})' and U^{-T} is lower triangular with positive
// diagonal, so by uniqueness chol_lower(C) = U^{-T} and
//     x = U^{-T} (U^{-1} b + z):
// one factorisation and two triangular solves give the reference's map exactly -- no explicit
// inverse and no second factorisation (5x fewer flops than inv + chol).
// hb = half bandwidth of A (A[i][j] = 0 for |i - j| > hb; n - 1 for a dense matrix).  The factor of a banded
// matrix has the same band, so B-spline Gram + tridiagonal penalty + diagonal shrinkage priors cost
// O(n hb^2) instead of O(n^3 / 3): at P = 20 (cubic, hb = 3) 180 instead of 2700 flops per block, at
// P = 400 (tensor basis, hb = 63) 13x fewer.  Skipped terms are exact zeros, so the result is the dense one.
// (implemented in host_linalg.cpp: baseline and AVX2+FMA clones of the same loops, selected at run time)
bool chol_upper_rev(int n, const double* A, double* U, int hb) { return bf_host::chol_upper_rev(n, A, U, hb); }
// x = U^{-T} (U^{-1} b + z)
void draw_from_rev_chol(int n, const double* U, const double* b, const double* z, double* x, double* w, int hb) {
  for (int i = n - 1; i >= 0; i--) {                 // U w = b
    const double* ui = U + (size_t)i * n;
    double s = b[i];
    const int ki = std::min(n - 1, i + hb);
    for (int k = i + 1; k <= ki; k++) s -= ui[k] * w[k];
    w[i] = s / ui[i];
  }
  for (int i = 0; i < n; i++) w[i] += z[i];
  for (int k = 0; k < n; k++) {                      // U' x = w  (column-oriented forward substitution)
    const double* uk = U + (size_t)k * n;
    const double xk = w[k] / uk[k];
    x[k] = xk;
    const int ik = std::min(n - 1, k + hb);
    for (int i = k + 1; i <= ik; i++) w[i] -= uk[i] * xk;
  }
}
// half bandwidth of a symmetric n x n matrix (exact zeros outside the band)
int half_bandwidth(int n, const double* A) {
  int hb = 0;
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++)
      if (A[(size_t)c * n + r] != 0.0 && std::abs(r - c) > hb) hb = std::abs(r - c);
  return hb;
}

// (normal / truncated-normal helpers: g_pnorm, g_qnorm, g_rtruncnorm_lo, g_dtruncnorm_lo_log in globals_core.cuh)

int sfail(const std::string& m) { return bf::set_error(m.c_str()); }

}  // namespace

struct bfmmm_sampler {
  bfmmm_engine* e = nullptr;
  bfmmm_hyper h{};
  int n = 0, K = 0, P = 0, M = 0, D = 0, q = 0;
  bool identity = false, ragged = false;
  int bw = 0;
  std::vector<double> zpre;     // normals of the Phi / nu block draws generated while the device was busy
  size_t zpre_pos = 0;          // (same streams, same order: the chain is unchanged)
  bool zpre_on = false;
  std::vector<std::vector<double>> pre_U;   // factors of the blocks' precisions computed in parallel ahead of the draws
  std::vector<char> pre_ok;
  std::vector<std::vector<double>> pre_Prec;  // per-thread scratch of prefactor_blocks
  std::vector<std::vector<double>> pre_prior; // scaled penalty matrices of the nu / eta blocks (kept: 1.3 MB each at P = 400)
  std::vector<double> prior_buf;
  int pre_next = -1;                        // next prefactored block (-1: none)
  uint32_t blk_purpose = 0;                 // stream (purpose, block index) of the next block draw's normals
  uint64_t blk_index = 0;
  int hbG = 0, hbP = 0;     // half bandwidths of the basis Gram and of the penalty matrix (block draws)
  const double* Hb = nullptr;   // ragged grids: pair cross-Gram band (set before the block draws)
  vecd Hb_own;
  int64_t n_total = 0, iteration = 0, last_accept = 0;
  int64_t tick = 0;        // monotone counter keying every random stream (tempered steps advance it too)
  bool in_tt = false;      // inside a tempered transition: sweeps do not advance `iteration`
  double sum_half_total = 0, n_points_total = 0;   // sum_i floor(n_i/2) and sum_i n_i over ALL shards
  HostRng rng;
  bfmmm_allreduce_fn allreduce = nullptr;
  void* allreduce_ctx = nullptr;
  // current values (Armadillo layouts)
  vecd nu, Phi, pi, delta, gamma, A, tau, eta, xi, tau_eta, delta_xi, gamma_xi, A_xi, Pmat, G;
  double sigma_sq = 1.0, alpha3 = 1.0, loglik = 0.0;
  vecd stats;     // host copy of the engine statistics buffer
  vecd tt_ssr, tt_sigma;   // per-slot trace of the last tempered transition
  double last_ssr = 0;     // SSR of the state the last sweep ended with
  double z_k1_us = -1, z_blocks_us = -1;   // Z proposal kernel alone / the host's Phi + nu block draws (microseconds)
  // The SSR after the chi step only feeds the reported log-likelihood, never the chain: its read-back
  // (and, on several GPUs, its all-reduce) is folded into the NEXT sweep's first exchange.
  bool ll_pending = false;
  double ll_sigma = 1.0;
  int64_t tt_accepts = 0, tt_total = 0;
  // wall-clock split of the sweeps run so far (seconds): host draws | waiting for the device (launch,
  // all-reduce, read-back) | pushing globals
  double t_host = 0, t_wait = 0, t_push = 0;
  // stored-sample recorder (BFMMM.h:1680-1746)
  struct Recorder {
    bool on = false;
    std::string dir;
    int r = 0, thin = 1, q = 0;
    std::vector<vecd> nu, Phi, pi, delta, gamma, A, tau, Z, chi, eta, xi, tau_eta, delta_xi, gamma_xi, A_xi;
    vecd sigma, alpha3;
  } rec;
  vecd work, Prec, C, Lc, rhs, v1, v2, zdraw;
  vecd fc_C, fc_Mt;          // block draws of one update: current coefficients (q x P) and M_b = sum_a S_ab c_a
  bool fc_valid = false;
  // ---- device-resident sweep (globals_kernels.cu): the chain's globals live in device memory, the sweep is a queue
  // of kernels, the host vectors above are a mirror refreshed on demand (dev_pull)
  struct Dev {
    bool on = false;          // the sampler runs its sweeps on the device
    bool host_stale = false;  // the device holds newer values than the host vectors
    bool dev_stale = true;    // the host vectors hold newer values than the device (set / host updates / restore)
    double* par = nullptr;    // [nu | Phi | pi (8) | alpha3 | sigma_sq | delta | gamma | A | tau] (device)
    double* cst = nullptr;    // [G | Pmat | L] (device)
    int* err = nullptr;       // device flag: a block precision was not positive definite
    long long* clk = nullptr; // BFMMM_DRAW_CLK=1: phase time stamps of the block-draw kernel (tuning)
    vecd h_par;               // host staging of `par`
    size_t o_nu = 0, o_Phi = 0, o_pi = 0, o_alpha3 = 0, o_sigma = 0, o_delta = 0, o_gamma = 0, o_A = 0, o_tau = 0, len = 0;
    size_t c_G = 0, c_P = 0, c_L = 0;
    int hbmax = 0;
    bf::EngineDevInfo info{};
    cudaStream_t side = nullptr;          // stream of the priors kernel
    cudaEvent_t ev_drawn = nullptr, ev_priors = nullptr, ev_stats = nullptr, ev_pi = nullptr;
    bool priors_pending = false, pi_pending = false;
    bool ll_from_ssr = false;             // the last sweep had no chi step: its log-likelihood comes from the SSR slot
  } dev;

  double& nu_(int k, int p) { return nu[(size_t)p * K + k]; }
  double& Phi_(int k, int p, int m) { return Phi[((size_t)m * P + p) * K + k]; }
  double& eta_(int p, int d, int k) { return eta[((size_t)k * D + d) * P + p]; }
  double& xi_(int k, int p, int d, int m) { return xi[(size_t)k * P * D * M + ((size_t)m * D + d) * P + p]; }
  double& gamma_(int k, int p, int m) { return gamma[((size_t)m * P + p) * K + k]; }
  double& gamma_xi_(int k, int p, int d, int m) { return gamma_xi[(size_t)k * P * D * M + ((size_t)m * D + d) * P + p]; }
  double& delta_(int k, int m) { return delta[(size_t)m * K + k]; }
  double& delta_xi_(int k, int m, int d) { return delta_xi[((size_t)d * M + m) * K + k]; }
  double& A_(int k, int i) { return A[(size_t)i * K + k]; }
  double& A_xi_(int k, int i, int d) { return A_xi[((size_t)d * 2 + i) * K + k]; }
  double& tau_eta_(int k, int d) { return tau_eta[(size_t)d * K + k]; }
  int feat(int k, int mm, int dd) const { return (k * (M + 1) + mm) * (1 + D) + dd; }
};

bfmmm_engine* bfmmm_sampler_engine(bfmmm_sampler* s) { return s ? s->e : nullptr; }

namespace {

// an update that asked for more injected draws than the tape held has produced NaNs: fail instead of returning them
int tape_check(bfmmm_sampler* s) {
  if (!s->rng.tape_underrun) return 0;
  s->rng.tape_underrun = false;
  return sfail("the tape of injected draws ran out (bfmmm_sampler_tape holds fewer values than the update consumes)");
}

GlobalsView host_view(bfmmm_sampler* s) {
  GlobalsView g;
  g.K = s->K; g.P = s->P; g.M = s->M; g.identity = s->identity ? 1 : 0; g.hbP = s->hbP;
  g.nu = s->nu.data(); g.Phi = s->Phi.data(); g.pi = s->pi.data(); g.delta = s->delta.data(); g.gamma = s->gamma.data();
  g.A = s->A.data(); g.tau = s->tau.data(); g.alpha3 = &s->alpha3;
  g.Pmat = s->Pmat.empty() ? nullptr : s->Pmat.data();
  g.h = s->h; g.n_total = (double)s->n_total;
  return g;
}
int dev_pull(bfmmm_sampler* s);
// a host-side update on a sampler whose chain lives on the device: fetch the current values first, and remember
// that the device copy is out of date afterwards
int dev_begin_host_update(bfmmm_sampler* s) {
  if (!s->dev.on) return 0;
  if (dev_pull(s)) return 1;
  s->dev.dev_stale = true;
  return 0;
}

// the chain's global parameters: what a rejected tempered transition restores (BFMMM.h:1631-1651)
struct ChainParams {
  vecd nu, Phi, pi, delta, gamma, A, tau, eta, xi, tau_eta, delta_xi, gamma_xi, A_xi;
  double sigma_sq, alpha3, loglik, last_ssr, ll_sigma;
  bool ll_pending;
  int64_t last_accept;
};
ChainParams save_params(const bfmmm_sampler* s) {
  return ChainParams{s->nu, s->Phi, s->pi, s->delta, s->gamma, s->A, s->tau, s->eta, s->xi, s->tau_eta, s->delta_xi, s->gamma_xi,
                     s->A_xi, s->sigma_sq, s->alpha3, s->loglik, s->last_ssr, s->ll_sigma, s->ll_pending, s->last_accept};
}
void restore_params(bfmmm_sampler* s, const ChainParams& c) {
  s->nu = c.nu; s->Phi = c.Phi; s->pi = c.pi; s->delta = c.delta; s->gamma = c.gamma; s->A = c.A; s->tau = c.tau;
  s->eta = c.eta; s->xi = c.xi; s->tau_eta = c.tau_eta; s->delta_xi = c.delta_xi; s->gamma_xi = c.gamma_xi; s->A_xi = c.A_xi;
  s->sigma_sq = c.sigma_sq; s->alpha3 = c.alpha3; s->loglik = c.loglik; s->last_ssr = c.last_ssr; s->ll_sigma = c.ll_sigma;
  s->ll_pending = c.ll_pending; s->last_accept = c.last_accept;
}

// coefficient vector of feature f (length P) read from / written to the sampler state
void get_coef(bfmmm_sampler* s, int k, int mm, int dd, double* out) {
  for (int p = 0; p < s->P; p++) {
    if (mm == 0 && dd == 0) out[p] = s->nu_(k, p);
    else if (mm == 0) out[p] = s->eta_(p, dd - 1, k);
    else if (dd == 0) out[p] = s->Phi_(k, p, mm - 1);
    else out[p] = s->xi_(k, p, dd - 1, mm - 1);
  }
}
void set_coef(bfmmm_sampler* s, int k, int mm, int dd, const double* in) {
  for (int p = 0; p < s->P; p++) {
    if (mm == 0 && dd == 0) s->nu_(k, p) = in[p];
    else if (mm == 0) s->eta_(p, dd - 1, k) = in[p];
    else if (dd == 0) s->Phi_(k, p, mm - 1) = in[p];
    else s->xi_(k, p, dd - 1, mm - 1) = in[p];
  }
}

// One Gaussian block draw from the sufficient statistics:
//   Prec = beta * S_aa * G / sigma^2 + Prior,   rhs = beta * (g_a - G * sum_{b != a} S_ab c_b) / sigma^2
//   C = pinv(Prec) symmetrised (nu, eta: UpdateNu.h:67-68) or inv(Prec) (Phi, xi: UpdatePhi.h:79)
//   draw = C rhs + chol_lower(C) z                                (arma::mvnrnd, UpdateNu.h:69)
// The precision matrices of the blocks of one update (all Phi blocks, or all nu blocks) depend on the
// statistics and the priors only -- not on the coefficients being drawn -- so for large P (the tensor-product
// basis of the high-dimensional model, P = 400) they are built and factorised on several host threads
// before the sequential draws, which then only need the right-hand side and two triangular solves.
struct PreBlock { int a; const double* prior_full; std::vector<double> prior_diag; };
// out = c * A on the band |r - c| <= hb (the entries outside it are zero in A and stay zero in out)
void scale_band(int n, int hb, double c, const double* A, double* out) {
  for (int col = 0; col < n; col++)
    for (int r = std::max(0, col - hb); r <= std::min(n - 1, col + hb); r++) out[(size_t)col * n + r] = c * A[(size_t)col * n + r];
}
constexpr int PREFACTOR_MIN_P = 96;
void prefactor_blocks(bfmmm_sampler* s, const std::vector<PreBlock>& blocks, const double* WtW, double beta) {
  const int P = s->P, q = s->q, n = (int)blocks.size();
  s->pre_U.resize(n); s->pre_ok.assign(n, 0);
  const double sc = beta / s->sigma_sq;
  const int hg = s->identity ? 0 : s->hbG;
  int nt = (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(std::min(n, 16), nt > 1 ? nt / 2 : 1));
  if ((int)s->pre_Prec.size() < nt) s->pre_Prec.resize(nt);
  auto work = [&](int t0, int stride) {
    std::vector<double>& Prec = s->pre_Prec[t0];     // only the band is written and read
    if (Prec.size() < (size_t)P * P) Prec.resize((size_t)P * P);
    for (int t = t0; t < n; t += stride) {
      const PreBlock& b = blocks[t];
      const double saa = WtW[(size_t)b.a * q + b.a];
      const int hbb = std::max(hg, b.prior_full ? s->hbP : 0);
      for (int c = 0; c < P; c++)
        for (int r = std::max(0, c - hbb); r <= std::min(P - 1, c + hbb); r++) {
          double g = s->identity ? (r == c ? 1.0 : 0.0) : s->G[(size_t)c * P + r];
          double pr = b.prior_full ? b.prior_full[(size_t)c * P + r] : (r == c ? b.prior_diag[r] : 0.0);
          Prec[(size_t)c * P + r] = sc * saa * g + pr;
        }
      if (s->pre_U[t].size() < (size_t)P * P) s->pre_U[t].resize((size_t)P * P);
      s->pre_ok[t] = chol_upper_rev(P, Prec.data(), s->pre_U[t].data(), hbb) ? 1 : 0;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work, t, nt);
  work(0, nt);
  for (auto& x : th) x.join();
}

// the P standard normals of the next block draw: element (block index, coefficient) of the update's stream -- or the
// values generated ahead while the device was busy (zpre), or the tape of injected draws
void block_normals(bfmmm_sampler* s, double* z) {
  const int P = s->P;
  if (s->zpre_on) { for (int p = 0; p < P; p++) z[p] = s->zpre[s->zpre_pos++]; }
  else {
    HostRngAdapter r{&s->rng};
    for (int p = 0; p < P; p++) { auto st = r.open(s->blk_purpose, (s->blk_index << 12) + (uint64_t)p); z[p] = st.normal(); }
  }
  s->blk_index++;
}
// Small P on a common basis: the band-storage factorisation and solves of globals_core.cuh, i.e. the statements the
// device-resident sweep runs (draw_blocks_kernel).  Returns -1 when the precision is not positive definite (the
// caller then takes the pseudo-inverse route of the reference).
int block_draw_band(bfmmm_sampler* s, int k, int mm, int dd, const double* WtW, const double* BtYW, double beta,
                    const double* prior_full, const double* prior_diag, const double* z) {
  const int P = s->P, q = s->q, a = s->feat(k, mm, dd);
  const double sc = beta / s->sigma_sq;
  const int hg = s->identity ? 0 : s->hbG, hb = std::max(hg, prior_full ? s->hbP : 0), ldb = hb + 1;
  s->v1.assign(P, 0.0); s->v2.resize(P); s->rhs.resize(P);
  // coefficients C (q x P) and M_b = sum_a' S_a'b c_a' (q x P) of the current update: built at its first block, then
  // kept current with a rank-one update per drawn block (the statements of draw_blocks_kernel)
  if (!s->fc_valid) {
    s->fc_C.assign((size_t)q * P, 0.0); s->fc_Mt.assign((size_t)q * P, 0.0);
    for (int kk = 0; kk < s->K; kk++)
      for (int m2 = 0; m2 <= s->M; m2++)
        for (int d2 = 0; d2 <= s->D; d2++) get_coef(s, kk, m2, d2, &s->fc_C[(size_t)s->feat(kk, m2, d2) * P]);
    for (int b = 0; b < q; b++)
      for (int f = 0; f < q; f++) {
        const double sfb = WtW[(size_t)f * q + b];
        if (sfb == 0.0) continue;
        const double* cf = &s->fc_C[(size_t)f * P];
        double* mb = &s->fc_Mt[(size_t)b * P];
        for (int p = 0; p < P; p++) mb[p] += sfb * cf[p];
      }
    s->fc_valid = true;
  }
  {
    const double saa0 = WtW[(size_t)a * q + a];
    const double* ca = &s->fc_C[(size_t)a * P];
    const double* ma = &s->fc_Mt[(size_t)a * P];
    for (int p = 0; p < P; p++) s->v1[p] = ma[p] - saa0 * ca[p];          // sum_{b != a} S_ab c_b
  }
  for (int r = 0; r < P; r++) {
    double gv = 0;
    if (s->identity) gv = s->v1[r];
    else for (int c = std::max(0, r - hg); c <= std::min(P - 1, r + hg); c++) gv += s->G[(size_t)r * P + c] * s->v1[c];
    s->rhs[r] = sc * (BtYW[(size_t)a * P + r] - gv);
  }
  const double saa = WtW[(size_t)a * q + a];
  s->work.resize((size_t)P * ldb + 3 * (size_t)P);
  double* A = s->work.data();
  double *rd = A + (size_t)P * ldb, *w = rd + P, *x = w + P;
  for (int i = 0; i < P; i++)
    for (int d = 0; d <= hb; d++) {
      const int c = i + d;
      double v = 0;
      if (c < P) {
        const double g = s->identity ? (d == 0 ? 1.0 : 0.0) : s->G[(size_t)c * P + i];
        const double pr = prior_full ? prior_full[(size_t)c * P + i] : (d == 0 ? prior_diag[i] : 0.0);
        v = sc * saa * g + pr;
      }
      A[(size_t)i * ldb + d] = v;
    }
  if (!band_chol_upper_rev_rd(P, hb, ldb, A, A, rd)) return -1;
  band_draw_rd(P, hb, ldb, A, rd, s->rhs.data(), z, x, w);
  set_coef(s, k, mm, dd, x);
  {
    double* ca = &s->fc_C[(size_t)a * P];
    for (int p = 0; p < P; p++) { w[p] = x[p] - ca[p]; ca[p] = x[p]; }
    for (int b = 0; b < q; b++) {
      const double sab = WtW[(size_t)a * q + b];
      if (sab == 0.0) continue;
      double* mb = &s->fc_Mt[(size_t)b * P];
      for (int p = 0; p < P; p++) mb[p] += sab * w[p];
    }
  }
  return 0;
}

int block_draw(bfmmm_sampler* s, int k, int mm, int dd, const double* WtW, const double* BtYW, double beta,
               const double* prior_full, const double* prior_diag) {
  const int P = s->P, q = s->q;
  const int a = s->feat(k, mm, dd);
  s->zdraw.resize(P);
  block_normals(s, s->zdraw.data());
  if (!s->ragged && P < PREFACTOR_MIN_P) {
    const int rc = block_draw_band(s, k, mm, dd, WtW, BtYW, beta, prior_full, prior_diag, s->zdraw.data());
    if (rc >= 0) return rc;
  }
  s->Prec.resize((size_t)P * P); s->C.resize((size_t)P * P); s->Lc.resize((size_t)P * P);
  s->rhs.resize(P); s->v1.resize(P); s->v2.resize(P);
  const double sc = beta / s->sigma_sq;
  if (s->ragged) {
    // H_ab = sum_i w_ia w_ib G_i, banded: Hb[pair][j*P + p] = H_ab[p-j][p]
    if (!s->Hb) return sfail("block draw: ragged sampler has no pair cross-Gram (bfmmm_sampler_set_hband)");
    const int bw = s->bw;
    auto pair_index = [&](int x, int y) { if (x > y) std::swap(x, y); return x * q - x * (x - 1) / 2 + (y - x); };
    std::fill(s->v1.begin(), s->v1.end(), 0.0);
    for (int kk = 0; kk < s->K; kk++)
      for (int m2 = 0; m2 <= s->M; m2++)
        for (int d2 = 0; d2 <= s->D; d2++) {
          int b = s->feat(kk, m2, d2);
          if (b == a) continue;
          const double* hb = s->Hb + (size_t)pair_index(a, b) * bw * P;
          get_coef(s, kk, m2, d2, s->v2.data());
          for (int p = 0; p < P; p++) {
            s->v1[p] += hb[p] * s->v2[p];
            for (int j = 1; j < bw && p - j >= 0; j++) {
              s->v1[p] += hb[(size_t)j * P + p] * s->v2[p - j];
              s->v1[p - j] += hb[(size_t)j * P + p] * s->v2[p];
            }
          }
        }
    for (int r = 0; r < P; r++) s->rhs[r] = sc * (BtYW[(size_t)a * P + r] - s->v1[r]);
    const double* haa = s->Hb + (size_t)pair_index(a, a) * bw * P;
    for (int c = 0; c < P; c++)
      for (int r = 0; r < P; r++) {
        int lo = r < c ? r : c, hi = r < c ? c : r;
        double g = (hi - lo < bw) ? haa[(size_t)(hi - lo) * P + hi] : 0.0;
        double pr = prior_full ? prior_full[(size_t)c * P + r] : (r == c ? prior_diag[r] : 0.0);
        s->Prec[(size_t)c * P + r] = sc * g + pr;
      }
  } else {
  // v1 = sum_{b != a} S_ab c_b
  std::fill(s->v1.begin(), s->v1.end(), 0.0);
  for (int kk = 0; kk < s->K; kk++)
    for (int m2 = 0; m2 <= s->M; m2++)
      for (int d2 = 0; d2 <= s->D; d2++) {
        int b = s->feat(kk, m2, d2);
        if (b == a) continue;
        double sab = WtW[(size_t)b * q + a];
        if (sab == 0.0) continue;
        get_coef(s, kk, m2, d2, s->v2.data());
        for (int p = 0; p < P; p++) s->v1[p] += sab * s->v2[p];
      }
  const double saa = WtW[(size_t)a * q + a];
  const int hg = s->identity ? 0 : s->hbG;
  for (int r = 0; r < P; r++) {
    double gv = 0;
    if (s->identity) gv = s->v1[r];
    else for (int c = std::max(0, r - hg); c <= std::min(P - 1, r + hg); c++) gv += s->G[(size_t)r * P + c] * s->v1[c];   // G symmetric: contiguous row
    s->rhs[r] = sc * (BtYW[(size_t)a * P + r] - gv);
  }
  const int hbb = std::max(hg, prior_full ? s->hbP : 0);
  const bool have_factor = s->pre_next >= 0 && s->pre_next < (int)s->pre_U.size() && s->pre_ok[s->pre_next];
  if (!have_factor)
  for (int c = 0; c < P; c++)
    for (int r = std::max(0, c - hbb); r <= std::min(P - 1, c + hbb); r++) {
      double g = s->identity ? (r == c ? 1.0 : 0.0) : s->G[(size_t)c * P + r];
      double pr = prior_full ? prior_full[(size_t)c * P + r] : (r == c ? prior_diag[r] : 0.0);
      s->Prec[(size_t)c * P + r] = sc * saa * g + pr;
    }
  }
  const int hb = s->ragged ? std::max(s->bw - 1, prior_full ? s->hbP : 0) : std::max(s->identity ? 0 : s->hbG, prior_full ? s->hbP : 0);
  s->work.resize((size_t)2 * P * P);
  for (int p = 0; p < P; p++) s->v2[p] = s->zdraw[p];
  double* U = s->work.data();
  const bool pre = !s->ragged && s->pre_next >= 0 && s->pre_next < (int)s->pre_U.size();
  const bool pre_good = pre && s->pre_ok[s->pre_next];
  if (pre_good) U = s->pre_U[s->pre_next].data();
  if (pre) s->pre_next++;
  if (pre_good || chol_upper_rev(P, s->Prec.data(), U, hb)) {
    draw_from_rev_chol(P, U, s->rhs.data(), s->v2.data(), s->v1.data(), s->work.data() + (size_t)P * P, hb);
  } else {
    // the dense fallback reads the whole matrix: fill what the band-limited build skipped
    if (!s->ragged)
      for (int c = 0; c < P; c++)
        for (int r = 0; r < P; r++)
          if (std::abs(r - c) > hb) s->Prec[(size_t)c * P + r] = 0.0;
    // singular precision: Moore-Penrose inverse, symmetrised, then mean + chol_lower(C) z as written
    // in the reference (UpdateNu.h:67-69)
    pinv_sym_jacobi(P, s->Prec.data(), s->C.data());
    for (int c = 0; c < P; c++)
      for (int r = c + 1; r < P; r++) {
        double t = (s->C[(size_t)c * P + r] + s->C[(size_t)r * P + c]) / 2;
        s->C[(size_t)c * P + r] = t; s->C[(size_t)r * P + c] = t;
      }
    if (!chol_rows(P, s->C.data(), s->Lc.data())) return sfail("block draw: covariance is not positive definite");
    for (int r = 0; r < P; r++) {
      const double* cr = s->C.data() + (size_t)r * P;      // symmetric: row r == column r
      const double* lr = s->Lc.data() + (size_t)r * P;
      double mean = 0, dev = 0;
      for (int c = 0; c < P; c++) mean += cr[c] * s->rhs[c];
      for (int c = 0; c <= r; c++) dev += lr[c] * s->v2[c];
      s->v1[r] = mean + dev;
    }
  }
  set_coef(s, k, mm, dd, s->v1.data());
  s->fc_valid = false;            // the band path's running sums no longer match the coefficients
  return 0;
}

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int push_globals(bfmmm_sampler* s) {
  struct T { bfmmm_sampler* s; double t0; ~T() { s->t_push += now_s() - t0; } } timer{s, now_s()};
  return bfmmm_set_globals(s->e, s->nu.data(), s->Phi.data(), s->D ? s->eta.data() : nullptr,
                           s->D ? s->xi.data() : nullptr, s->sigma_sq);
}

// Sums the statistics over the shards (hook) and reads them back.  only_ssr: exchange just the SSR slot
// (8 bytes).  Every slot must be summed exactly once after the kernel that wrote it: the full exchange
// follows the Z / statistics kernels (and carries the previous sweep's post-chi SSR), the one-slot
// exchange follows the SSR kernel.
int reduce_and_read(bfmmm_sampler* s, bool only_ssr = false) {
  struct T { bfmmm_sampler* s; double t0; ~T() { s->t_wait += now_s() - t0; } } timer{s, now_s()};
  double* dev = nullptr; int64_t len = 0;
  // the statistics pass may have summed the whole buffer over the shards in its own epilogue (peer-memory exchange)
  const bool exchanged = !only_ssr && bfmmm_stats_exchanged(s->e);
  if (exchanged || !s->allreduce) len = bfmmm_stats_len(s->e);      // nobody touches the buffer: the epilogue's host copy stands
  else if (bfmmm_stats_buffer_dev(s->e, &dev, &len)) return 1;
  if (s->allreduce && !exchanged) {
    int rc = only_ssr ? s->allreduce(s->allreduce_ctx, dev + s->K + 1, 1, bfmmm_stream(s->e))
                      : s->allreduce(s->allreduce_ctx, dev, len, bfmmm_stream(s->e));
    if (rc) return sfail("all-reduce hook failed");
  }
  s->stats.resize(len);
  return bfmmm_read_stats(s->e, s->stats.data(), only_ssr ? (int64_t)(s->K + 3) : len);
}
void set_loglik(bfmmm_sampler* s, double ssr_ll, double sigma_sq) {
  // calcLikelihood (CalculateLikelihood.h:19-44; MV :137-159 with floor(P/2))
  if (s->identity)
    s->loglik = -((double)s->n_total * (double)(s->P / 2)) * std::log(2 * 3.14159265358979323846 * sigma_sq) -
                ssr_ll / (2 * sigma_sq);
  else
    s->loglik = -s->n_points_total * (0.918938533204672741780329736406 + 0.5 * std::log(sigma_sq)) -
                ssr_ll / (2 * sigma_sq);
}
// completes a deferred log-likelihood outside a sweep (bfmmm_sampler_get, tempered transitions)
int flush_loglik(bfmmm_sampler* s) {
  if (!s->ll_pending || !s->e) return 0;
  if (s->dev.on) {                        // sigma^2 of the sweep is on the device
    if (dev_pull(s)) return 1;
    s->ll_sigma = s->sigma_sq;
  }
  struct T { bfmmm_sampler* s; double t0; ~T() { s->t_wait += now_s() - t0; } } timer{s, now_s()};
  double* dev = nullptr; int64_t len = 0;
  if (bfmmm_stats_buffer_dev(s->e, &dev, &len)) return 1;
  if (s->allreduce && s->allreduce(s->allreduce_ctx, dev + s->K + 2, 1, bfmmm_stream(s->e))) return sfail("all-reduce hook failed");
  vecd tmp(s->K + 3);
  if (bfmmm_read_stats(s->e, tmp.data(), s->K + 3)) return 1;
  s->last_ssr = tmp[s->K + 2];
  set_loglik(s, s->last_ssr, s->ll_sigma);
  s->ll_pending = false;
  // the slot now holds the global sum on every rank: zero it so the next full exchange does not add it again
  return bfmmm_clear_ssr_after(s->e);
}
int record_iteration_fwd(bfmmm_sampler* s);
const double* st_slz(bfmmm_sampler* s) { return s->stats.data(); }
double st_acc(bfmmm_sampler* s) { return s->stats[s->K]; }
double st_ssr(bfmmm_sampler* s) { return s->stats[s->K + 1]; }
double st_ssr_after(bfmmm_sampler* s) { return s->stats[s->K + 2]; }
const double* st_wtw(bfmmm_sampler* s) { return s->stats.data() + s->K + 3; }
const double* st_btyw(bfmmm_sampler* s) { return s->stats.data() + s->K + 3 + (size_t)s->q * s->q; }
const double* st_hb(bfmmm_sampler* s) { return s->stats.data() + s->K + 3 + (size_t)s->q * s->q + (size_t)s->P * s->q; }


// ================================================================= device-resident sweep
// The chain's globals live in device memory (s->dev.par); a sweep is a queue of kernels on the engine's stream -- the
// pass kernels, the Gaussian block draws, the sigma^2 / pi / alpha_3 draws -- plus the shrinkage priors on a side
// stream (they feed the next sweep's block draws and overlap this sweep's SSR and chi passes).  Nothing is read back
// unless the caller asks (bfmmm_sampler_get, the recorder, a tempered transition).
#define CUS(x)                                                                          \
  do {                                                                                  \
    cudaError_t _e = (x);                                                               \
    if (_e != cudaSuccess) return sfail(std::string(#x) + ": " + cudaGetErrorString(_e)); \
  } while (0)

constexpr size_t DEV_MAX_BAND_DOUBLES = 16384;      // nb * P * (hb + 2) doubles of factors the draw kernel keeps in shared memory

bool dev_eligible(const bfmmm_sampler* s) {
  if (!s->e || s->D > 0 || s->ragged || s->K > 8) return false;
  // Opt-in (BFMMM_DEVICE_GLOBALS=1): measured on B200 the single-block draw kernel is latency bound (52 us against the
  // 31 us the host spends on the same draws), so the host-drawn globals stay the default.
  if (!std::getenv("BFMMM_DEVICE_GLOBALS") || std::getenv("BFMMM_HOST_GLOBALS")) return false;
  const int hb = std::max(s->identity ? 0 : s->hbG, s->identity ? 0 : s->hbP);
  const size_t nb = (size_t)s->K * (s->M + 1);
  if (hb > 0 && s->P > 30) return false;      // banded precisions: one warp per block, a row of T_a in a lane's registers
  return nb * s->P * (hb + 5) <= DEV_MAX_BAND_DOUBLES;
}
int dev_init(bfmmm_sampler* s) {
  auto& d = s->dev;
  if (bfmmm_engine_devinfo(s->e, &d.info)) return 1;
  CUS(cudaSetDevice(d.info.device));
  const size_t K = s->K, P = s->P, M = s->M;
  size_t o = 0;
  d.o_nu = o; o += K * P;
  d.o_Phi = o; o += K * P * M;
  d.o_pi = o; o += 8;                 // [pi (8) | alpha3 | sigma_sq] is what the Z kernel reads
  d.o_alpha3 = o; o += 1;
  d.o_sigma = o; o += 1;
  d.o_delta = o; o += K * M;
  d.o_gamma = o; o += K * P * M;
  d.o_A = o; o += 2 * K;
  d.o_tau = o; o += K;
  d.len = o;
  d.h_par.assign(d.len, 0.0);
  CUS(cudaMalloc(&d.par, d.len * 8));
  CUS(cudaMalloc(&d.err, 4));
  CUS(cudaMemset(d.err, 0, 4));
  if (std::getenv("BFMMM_DRAW_CLK")) { CUS(cudaMalloc(&d.clk, 16 * 8)); CUS(cudaMemset(d.clk, 0, 16 * 8)); }
  const size_t Pc = d.info.Pc;
  d.c_G = 0; d.c_P = P * P; d.c_L = 2 * P * P;
  vecd cst(2 * P * P + P * Pc, 0.0);
  std::copy(s->G.begin(), s->G.end(), cst.begin() + d.c_G);
  if (!s->Pmat.empty()) std::copy(s->Pmat.begin(), s->Pmat.end(), cst.begin() + d.c_P);
  if (!s->identity) std::copy(d.info.L_host, d.info.L_host + P * Pc, cst.begin() + d.c_L);
  CUS(cudaMalloc(&d.cst, cst.size() * 8));
  CUS(cudaMemcpy(d.cst, cst.data(), cst.size() * 8, cudaMemcpyHostToDevice));
  d.hbmax = std::max(s->identity ? 0 : s->hbG, s->identity ? 0 : s->hbP);
  CUS(cudaStreamCreateWithFlags(&d.side, cudaStreamNonBlocking));
  CUS(cudaEventCreateWithFlags(&d.ev_drawn, cudaEventDisableTiming));
  CUS(cudaEventCreateWithFlags(&d.ev_priors, cudaEventDisableTiming));
  CUS(cudaEventCreateWithFlags(&d.ev_stats, cudaEventDisableTiming));
  CUS(cudaEventCreateWithFlags(&d.ev_pi, cudaEventDisableTiming));
  d.on = true; d.dev_stale = true; d.host_stale = false;
  return 0;
}
void dev_free(bfmmm_sampler* s) {
  auto& d = s->dev;
  if (!d.par && !d.cst) return;
  cudaSetDevice(d.info.device);
  if (d.side) { cudaStreamSynchronize(d.side); cudaStreamDestroy(d.side); }
  if (d.clk) {
    long long c[16];
    if (cudaMemcpy(c, d.clk, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess) {
      std::fprintf(stderr, "draw_blocks_kernel phases (cycles):");
      for (int i = 1; i < 8; i++) std::fprintf(stderr, " %lld", c[i] - c[i - 1]);
      std::fprintf(stderr, "\n");
    }
    cudaFree(d.clk);
  }
  if (d.ev_drawn) cudaEventDestroy(d.ev_drawn);
  if (d.ev_priors) cudaEventDestroy(d.ev_priors);
  if (d.ev_stats) cudaEventDestroy(d.ev_stats);
  if (d.ev_pi) cudaEventDestroy(d.ev_pi);
  cudaFree(d.par); cudaFree(d.cst); cudaFree(d.err);
  d = bfmmm_sampler::Dev();
}
GlobalsView dev_view(bfmmm_sampler* s) {
  auto& d = s->dev;
  GlobalsView g = host_view(s);
  g.nu = d.par + d.o_nu; g.Phi = d.par + d.o_Phi; g.pi = d.par + d.o_pi; g.delta = d.par + d.o_delta;
  g.gamma = d.par + d.o_gamma; g.A = d.par + d.o_A; g.tau = d.par + d.o_tau; g.alpha3 = d.par + d.o_alpha3;
  g.Pmat = s->identity ? nullptr : d.cst + d.c_P;
  return g;
}
// host vectors -> device (after bfmmm_sampler_set, a host-side update or a restored tempered transition)
int dev_push(bfmmm_sampler* s) {
  auto& d = s->dev;
  if (!d.on || !d.dev_stale) return 0;
  CUS(cudaSetDevice(d.info.device));
  // the staging buffer may still be travelling from an earlier push
  CUS(cudaStreamSynchronize(d.info.stream));
  double* h = d.h_par.data();
  std::copy(s->nu.begin(), s->nu.end(), h + d.o_nu);
  std::copy(s->Phi.begin(), s->Phi.end(), h + d.o_Phi);
  std::fill(h + d.o_pi, h + d.o_pi + 8, 0.0);
  std::copy(s->pi.begin(), s->pi.end(), h + d.o_pi);
  h[d.o_alpha3] = s->alpha3; h[d.o_sigma] = s->sigma_sq;
  std::copy(s->delta.begin(), s->delta.end(), h + d.o_delta);
  std::copy(s->gamma.begin(), s->gamma.end(), h + d.o_gamma);
  std::copy(s->A.begin(), s->A.end(), h + d.o_A);
  std::copy(s->tau.begin(), s->tau.end(), h + d.o_tau);
  if (d.priors_pending) { CUS(cudaStreamWaitEvent(d.info.stream, d.ev_priors, 0)); d.priors_pending = false; }
  if (d.pi_pending) { CUS(cudaStreamWaitEvent(d.info.stream, d.ev_pi, 0)); d.pi_pending = false; }
  CUS(cudaMemcpyAsync(d.par, h, d.len * 8, cudaMemcpyHostToDevice, d.info.stream));
  if (push_globals(s)) return 1;             // the whitened coefficients of the pass kernels
  d.dev_stale = false;
  return 0;
}
// device -> host vectors (get, recorder, tempered transitions, host-side updates)
int dev_pull(bfmmm_sampler* s) {
  auto& d = s->dev;
  if (!d.on || !d.host_stale) return 0;
  CUS(cudaSetDevice(d.info.device));
  CUS(cudaStreamSynchronize(d.side));
  CUS(cudaStreamSynchronize(d.info.stream));
  int err = 0;
  CUS(cudaMemcpy(&err, d.err, 4, cudaMemcpyDeviceToHost));
  if (err) return sfail("device block draw: a precision matrix is not positive definite (BFMMM_HOST_GLOBALS=1 selects the host draws with the reference's pseudo-inverse route)");
  CUS(cudaMemcpy(d.h_par.data(), d.par, d.len * 8, cudaMemcpyDeviceToHost));
  const double* h = d.h_par.data();
  std::copy(h + d.o_nu, h + d.o_nu + s->nu.size(), s->nu.begin());
  std::copy(h + d.o_Phi, h + d.o_Phi + s->Phi.size(), s->Phi.begin());
  std::copy(h + d.o_pi, h + d.o_pi + s->pi.size(), s->pi.begin());
  s->alpha3 = h[d.o_alpha3]; s->sigma_sq = h[d.o_sigma];
  std::copy(h + d.o_delta, h + d.o_delta + s->delta.size(), s->delta.begin());
  std::copy(h + d.o_gamma, h + d.o_gamma + s->gamma.size(), s->gamma.begin());
  std::copy(h + d.o_A, h + d.o_A + s->A.size(), s->A.begin());
  std::copy(h + d.o_tau, h + d.o_tau + s->tau.size(), s->tau.begin());
  if (bfmmm_set_sigma(s->e, s->sigma_sq)) return 1;
  // acceptance count and SSR of the last sweep (local shard unless the exchange already summed them)
  vecd st(s->K + 3);
  if (bfmmm_read_stats(s->e, st.data(), s->K + 3)) return 1;
  s->last_accept = (int64_t)std::llround(st[s->K]);
  if ((int64_t)s->stats.size() < s->K + 3) s->stats.resize(s->K + 3);
  std::copy(st.begin(), st.end(), s->stats.begin());
  d.host_stale = false;
  if (d.ll_from_ssr) {                 // Nu_Z sweeps: calcLikelihood on the SSR of the sigma^2 step (already summed over shards)
    s->last_ssr = st[s->K + 1];
    set_loglik(s, s->last_ssr, s->sigma_sq);
  }
  return 0;
}

int sampler_step_device(bfmmm_sampler* s, int sweep, double beta) {
  auto& d = s->dev;
  bfmmm_engine* e = s->e;
  const bool do_z = (sweep == BFMMM_SWEEP_NU_Z || sweep == BFMMM_SWEEP_FULL);
  const bool do_phi = (sweep == BFMMM_SWEEP_THETA || sweep == BFMMM_SWEEP_FULL);
  const bool do_nu = do_z, do_chi = do_phi;
  const bool tempered = s->in_tt || beta != 1.0;
  s->rng.iteration = (uint64_t)s->tick;
  CUS(cudaSetDevice(d.info.device));
  if (dev_push(s)) return 1;
  if (bfmmm_seed(e, s->rng.key, (uint64_t)s->tick)) return 1;
  cudaStream_t st = d.info.stream;
  const double* zpar = d.par + d.o_pi;
  double* sigma_dev = d.par + d.o_sigma;
  const StreamRng rng{s->rng.key, (uint64_t)s->tick};
  if (d.pi_pending) { CUS(cudaStreamWaitEvent(st, d.ev_pi, 0)); d.pi_pending = false; }     // last sweep's pi, alpha_3
  if (do_z && bfmmm_update_z_async_p(e, s->h.a_Z_PM, beta, zpar)) return 1;          // updateZ_PM
  if (bfmmm_suffstats_async(e)) return 1;
  if (s->allreduce && !bfmmm_stats_exchanged(e) && s->allreduce(s->allreduce_ctx, d.info.stats, d.info.stats_len, st))
    return sfail("all-reduce hook failed");
  CUS(cudaEventRecord(d.ev_stats, st));
  // updatePhi, updateNu: the blocks need last sweep's delta, gamma, tau (side stream)
  if (d.priors_pending) { CUS(cudaStreamWaitEvent(st, d.ev_priors, 0)); d.priors_pending = false; }
  DrawArgs da;
  da.g = dev_view(s);
  da.Pc = d.info.Pc; da.P4 = d.info.P4; da.QS = d.info.QS; da.q = s->q;
  da.hbG = s->identity ? 0 : s->hbG; da.hbL = d.info.hbL; da.hbmax = d.hbmax;
  da.G = s->identity ? nullptr : d.cst + d.c_G; da.L = s->identity ? nullptr : d.cst + d.c_L;
  da.stats = d.info.stats; da.glob = d.info.glob; da.sigma_dev = sigma_dev; da.beta = beta;
  da.do_phi = do_phi ? 1 : 0; da.do_nu = do_nu ? 1 : 0; da.rng = rng; da.err = d.err;
  da.clk = d.clk;
  if (launch_draw_blocks(da, st)) return sfail("draw_blocks kernel launch failed");
  bfmmm_moments_invalidate(e);          // the kernel writes the staged globals
  CUS(cudaEventRecord(d.ev_drawn, st));
  if (bfmmm_ssr_async(e)) return 1;                                                 // updateSigma's data pass, new globals
  if (s->allreduce && s->allreduce(s->allreduce_ctx, d.info.stats + s->K + 1, 1, st)) return sfail("all-reduce hook failed");
  SigmaPiArgs sa;
  sa.g = da.g; sa.stats = d.info.stats; sa.sigma_dev = sigma_dev;
  sa.shape = tempered ? (beta * s->n_points_total) / 2 + s->h.alpha_0 : s->sum_half_total + s->h.alpha_0;
  sa.scale_ssr = tempered ? beta / 2 : 0.5;
  sa.do_pi = do_z ? 1 : 0; sa.rng = rng;
  if (launch_sigma(sa, st)) return sfail("sigma kernel launch failed");              // updateSigma
  if (do_chi && bfmmm_update_chi_async_p(e, beta, sigma_dev)) return 1;             // updateChi (+ the SSR calcLikelihood needs)
  // side stream: updatePi_PM, updateAlpha3 (they only need the reduced sum_i log Z_ik), then, behind the block draws,
  // updateDelta, updateA, updateGamma, updateTau
  if (do_z) {
    CUS(cudaStreamWaitEvent(d.side, d.ev_stats, 0));
    if (launch_pi_alpha(sa, d.side)) return sfail("pi / alpha_3 kernel launch failed");
    CUS(cudaEventRecord(d.ev_pi, d.side));
    d.pi_pending = true;
  }
  CUS(cudaStreamWaitEvent(d.side, d.ev_drawn, 0));
  PriorsArgs pa;
  pa.g = da.g; pa.do_phi = do_phi ? 1 : 0; pa.rng = rng;
  if (launch_priors(pa, d.side)) return sfail("priors kernel launch failed");
  CUS(cudaEventRecord(d.ev_priors, d.side));
  d.priors_pending = true;
  d.host_stale = true;
  s->ll_pending = do_chi; s->ll_sigma = 0.0;                 // sigma^2 is on the device: flush_loglik pulls it
  d.ll_from_ssr = !do_chi;
  s->tick++;
  if (s->in_tt || s->rec.on) {
    if (dev_pull(s)) return 1;
    if (do_chi && flush_loglik(s)) return 1;

  }
  if (!s->in_tt) {
    if (s->rec.on && record_iteration_fwd(s)) return 1;
    s->iteration++;
  }
  return 0;
}

}  // namespace

extern "C" {

void bfmmm_hyper_defaults(bfmmm_hyper* h, int theta_est_defaults) {
  for (int k = 0; k < 8; k++) h->c[k] = 10.0;
  h->b = 10; h->nu_1 = 3;
  if (theta_est_defaults) { h->alpha1l = 2; h->alpha2l = 3; h->beta1l = 2; h->beta2l = 2; }   // UserFunctions.cpp:700-703
  else { h->alpha1l = 1; h->alpha2l = 2; h->beta1l = 1; h->beta2l = 1; }                      // :179-182
  h->a_Z_PM = 10000; h->a_pi_PM = 1000; h->var_alpha3 = 0.05; h->var_epsilon1 = 1; h->var_epsilon2 = 1;
  h->alpha_nu = 10; h->beta_nu = 1; h->alpha_eta = 10; h->beta_eta = 1; h->alpha_0 = 1; h->beta_0 = 1;
}

static bfmmm_sampler* make_sampler(const int32_t* dims, const bfmmm_hyper* h, int64_t n_total, uint64_t seed) {
  bfmmm_sampler* s = new bfmmm_sampler();
  s->h = *h;
  s->n = dims[0]; s->K = dims[1]; s->P = dims[2]; s->M = dims[3]; s->D = dims[4];
  s->identity = dims[5] == BFMMM_MULTIVARIATE;
  s->ragged = dims[6] != 0; s->bw = dims[7];
  s->q = s->K * (1 + s->D) * (1 + s->M);
  s->n_total = n_total > 0 ? n_total : s->n;
  s->rng.key = seed;
  const int K = s->K, P = s->P, M = s->M, D = s->D;
  s->nu.assign((size_t)K * P, 0.0); s->Phi.assign((size_t)K * P * M, 0.0);
  s->pi.assign(K, 1.0 / K); s->delta.assign((size_t)K * M, 1.0); s->gamma.assign((size_t)K * P * M, 1.0);
  s->A.assign((size_t)K * 2, 1.0); s->tau.assign(K, 1.0);
  if (D) {
    s->eta.assign((size_t)P * D * K, 0.0); s->xi.assign((size_t)K * P * D * M, 0.0);
    s->tau_eta.assign((size_t)K * D, 1.0); s->delta_xi.assign((size_t)K * M * D, 1.0);
    s->gamma_xi.assign((size_t)K * P * D * M, 1.0); s->A_xi.assign((size_t)K * 2 * D, 1.0);
  }
  s->G.assign((size_t)P * P, 0.0);
  return s;
}

int bfmmm_sampler_create(bfmmm_engine* e, const bfmmm_hyper* h, int64_t n_total, const double* Pmat,
                         uint64_t seed, bfmmm_sampler** out) {
  if (!e || !h || !out) return sfail("bfmmm_sampler_create: null argument");
  int32_t dims[8];
  if (bfmmm_engine_dims(e, dims)) return 1;
  if (dims[5] != BFMMM_MULTIVARIATE && !Pmat) return sfail("bfmmm_sampler_create: the functional model needs the penalty matrix P");
  bfmmm_sampler* s = make_sampler(dims, h, n_total, seed);
  s->e = e;
  if (!s->ragged) bfmmm_get_gram(e, s->G.data());
  if (Pmat) s->Pmat.assign(Pmat, Pmat + (size_t)s->P * s->P);
  s->hbG = half_bandwidth(s->P, s->G.data());
  s->hbP = Pmat ? half_bandwidth(s->P, s->Pmat.data()) : 0;
  double sum_half = 0, npts = 0;
  bfmmm_counts(e, &sum_half, &npts);
  // per-shard counts -> whole data set (common grid: every function has the same n_i)
  double ratio = (double)s->n_total / (double)s->n;   // exact on a common grid; ragged multi-shard runs
                                                       // set the totals with bfmmm_sampler_set_counts
  s->n_points_total = npts * ratio;
  s->sum_half_total = s->identity ? (double)(((int64_t)s->n_total * s->P) / 2) : sum_half * ratio;
  if (dev_eligible(s) && dev_init(s)) { delete s; return 1; }
  *out = s;
  return 0;
}

// sampler without an engine: only the bfmmm_host_update_* functions may be called on it (CPU tests)
int bfmmm_sampler_create_detached(const int32_t* dims, const bfmmm_hyper* h, int64_t n_total, const double* Pmat,
                                  const double* G, double sum_half_total, double n_points_total, uint64_t seed,
                                  bfmmm_sampler** out) {
  if (!dims || !h || !out) return sfail("bfmmm_sampler_create_detached: null argument");
  bfmmm_sampler* s = make_sampler(dims, h, n_total, seed);
  if (G) s->G.assign(G, G + (size_t)s->P * s->P);
  if (Pmat) s->Pmat.assign(Pmat, Pmat + (size_t)s->P * s->P);
  s->hbG = half_bandwidth(s->P, s->G.data());
  s->hbP = Pmat ? half_bandwidth(s->P, s->Pmat.data()) : 0;
  s->sum_half_total = sum_half_total; s->n_points_total = n_points_total;
  *out = s;
  return 0;
}
void bfmmm_sampler_destroy(bfmmm_sampler* s) {
  if (s) dev_free(s);
  // an exchange fused into the engine's passes points into the hook's context (mailboxes, sequence counter), which the
  // caller frees after the sampler: the engine must not outlive it with that wiring
  if (s && s->e) bfmmm_engine_set_exchange(s->e, nullptr, 0, 1, 0, nullptr);
  delete s;
}
// 1 when the sampler's sweeps run device-resident (globals_kernels.cu), 0 when the host draws the globals
int bfmmm_sampler_device_resident(bfmmm_sampler* s) { return s && s->dev.on ? 1 : 0; }
// ragged grids: pair cross-Gram band for the bfmmm_host_update_{phi,nu,eta,xi} calls that follow
// (copied); bfmmm_sampler_step sets it from the device statistics itself
int bfmmm_sampler_set_hband(bfmmm_sampler* s, const double* Hband) {
  if (!s || !s->ragged) return sfail("bfmmm_sampler_set_hband: not a ragged-grid sampler");
  size_t len = (size_t)(s->q * (s->q + 1) / 2) * s->bw * s->P;
  s->Hb_own.assign(Hband, Hband + len);
  s->Hb = s->Hb_own.data();
  return 0;
}
// totals over ALL shards of sum_i floor(n_i/2) and sum_i n_i (ragged multi-GPU runs all-reduce the
// per-shard bfmmm_counts once and set them here)
int bfmmm_sampler_set_counts(bfmmm_sampler* s, double sum_half_total, double n_points_total) {
  if (!s) return sfail("null sampler");
  s->sum_half_total = sum_half_total; s->n_points_total = n_points_total;
  return 0;
}
int bfmmm_sampler_set_allreduce(bfmmm_sampler* s, bfmmm_allreduce_fn fn, void* ctx) {
  if (!s) return sfail("null sampler");
  // a new hook replaces the fused peer-memory exchange of an earlier one (bfmmm_sampler_enable_p2p re-installs its own)
  if (s->e && bfmmm_engine_set_exchange(s->e, nullptr, 0, 1, 0, nullptr)) return 1;
  s->allreduce = fn; s->allreduce_ctx = ctx;
  return 0;
}

#define CP_IN(dst, src) if (src) std::copy(src, src + (dst).size(), (dst).begin())
#define CP_OUT(dst, src) if (dst) std::copy((src).begin(), (src).end(), dst)
int bfmmm_sampler_set(bfmmm_sampler* s, const double* nu, const double* Phi, const double* sigma_sq,
                      const double* pi, const double* alpha3, const double* delta, const double* gamma,
                      const double* A, const double* tau) {
  if (!s) return sfail("null sampler");
  if (dev_begin_host_update(s)) return 1;
  CP_IN(s->nu, nu); CP_IN(s->Phi, Phi); CP_IN(s->pi, pi); CP_IN(s->delta, delta); CP_IN(s->gamma, gamma);
  CP_IN(s->A, A); CP_IN(s->tau, tau);
  if (sigma_sq) s->sigma_sq = *sigma_sq;
  if (alpha3) s->alpha3 = *alpha3;
  return 0;
}
int bfmmm_sampler_get(bfmmm_sampler* s, double* nu, double* Phi, double* sigma_sq, double* pi,
                      double* alpha3, double* delta, double* gamma, double* A, double* tau, double* loglik) {
  if (!s) return sfail("null sampler");
  if (dev_pull(s)) return 1;
  if (loglik && flush_loglik(s)) return 1;
  CP_OUT(nu, s->nu); CP_OUT(Phi, s->Phi); CP_OUT(pi, s->pi); CP_OUT(delta, s->delta); CP_OUT(gamma, s->gamma);
  CP_OUT(A, s->A); CP_OUT(tau, s->tau);
  if (sigma_sq) *sigma_sq = s->sigma_sq;
  if (alpha3) *alpha3 = s->alpha3;
  if (loglik) *loglik = s->loglik;
  return 0;
}
int bfmmm_sampler_set_cov(bfmmm_sampler* s, const double* eta, const double* xi, const double* tau_eta,
                          const double* delta_xi, const double* gamma_xi, const double* A_xi) {
  if (!s || !s->D) return sfail("sampler has no covariates");
  CP_IN(s->eta, eta); CP_IN(s->xi, xi); CP_IN(s->tau_eta, tau_eta); CP_IN(s->delta_xi, delta_xi);
  CP_IN(s->gamma_xi, gamma_xi); CP_IN(s->A_xi, A_xi);
  return 0;
}
int bfmmm_sampler_get_cov(bfmmm_sampler* s, double* eta, double* xi, double* tau_eta, double* delta_xi,
                          double* gamma_xi, double* A_xi) {
  if (!s || !s->D) return sfail("sampler has no covariates");
  CP_OUT(eta, s->eta); CP_OUT(xi, s->xi); CP_OUT(tau_eta, s->tau_eta); CP_OUT(delta_xi, s->delta_xi);
  CP_OUT(gamma_xi, s->gamma_xi); CP_OUT(A_xi, s->A_xi);
  return 0;
}
int bfmmm_sampler_tape(bfmmm_sampler* s, const double* values, int64_t n) {
  if (!s) return sfail("null sampler");
  if (s->dev.on) {                       // injected draws follow the reference's sequential call order: host path
    if (dev_pull(s)) return 1;
    s->dev.on = false;
  }
  s->rng.use_tape = true;
  for (int64_t i = 0; i < n; i++) s->rng.tape.push_back(values[i]);
  return 0;
}
// positions the global random streams (keyed by seed, tick, purpose): the single bfmmm_host_update_* calls draw from
// the stream of the current tick, so a caller iterating one update advances it between calls
int bfmmm_sampler_set_tick(bfmmm_sampler* s, int64_t tick) {
  if (!s) return sfail("null sampler");
  s->tick = tick; s->rng.iteration = (uint64_t)tick;
  return 0;
}
int64_t bfmmm_sampler_tape_left(bfmmm_sampler* s) { return s ? (int64_t)s->rng.tape.size() : -1; }
int64_t bfmmm_sampler_iteration(bfmmm_sampler* s) { return s ? s->iteration : -1; }
int64_t bfmmm_sampler_last_accept(bfmmm_sampler* s) {
  if (!s || dev_pull(s)) return -1;
  return s->last_accept;
}

// ================================================================= host-side updates
// The prior updates below run the statements of globals_core.cuh (shared with the device-resident sweep):
// updatePi_PM (UpdatePi.h:84-116; lpdf_pi_PM :39-53 with sum_i log Z_ik from the device)
int bfmmm_host_update_pi(bfmmm_sampler* s, const double* slz) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  core_update_pi(g, r, slz);
  return tape_check(s);
}

// updateAlpha3 (UpdateAlpha3.h:36-63, lpdf_alpha3 :10-26)
int bfmmm_host_update_alpha3(bfmmm_sampler* s, const double* slz) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  core_update_alpha3(g, r, slz);
  return tape_check(s);
}

// updateTau (UpdateTau.h:18-40) / updateTauMV (:47-68): note the integer division nu.n_cols / 2
int bfmmm_host_update_tau(bfmmm_sampler* s) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  for (int k = 0; k < s->K; k++) core_update_tau_k(g, r, k);
  return tape_check(s);
}

// updateTauEta (UpdateTau.h:75-99) / updateTauEtaMV (:106-128)
int bfmmm_host_update_tau_eta(bfmmm_sampler* s) {
  const int K = s->K, P = s->P, D = s->D;
  s->rng.open(HP_TAU_ETA);
  for (int j = 0; j < K; j++)
    for (int d = 0; d < D; d++) {
      double a = s->h.alpha_eta + (double)(P / 2);
      double quad = 0;
      for (int r = 0; r < P; r++) {
        double pr = 0;
        if (s->identity) pr = s->eta_(r, d, j);
        else for (int c = std::max(0, r - s->hbP); c <= std::min(P - 1, r + s->hbP); c++) pr += s->Pmat[(size_t)r * P + c] * s->eta_(c, d, j);
        quad += s->eta_(r, d, j) * pr;
      }
      double b = s->h.beta_eta + 0.5 * quad;
      double g = (1 / b) * s->rng.gamma(a);
      s->tau_eta_(j, d) = s->identity ? 1 / g : g;
    }
  return tape_check(s);
}

// updateDelta (UpdateDelta.h:17-66): multiplicative gamma process shrinkage
int bfmmm_host_update_delta(bfmmm_sampler* s) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  for (int k = 0; k < s->K; k++) core_update_delta_k(g, r, k);
  return tape_check(s);
}

// updateDeltaXi (UpdateDelta.h:76-125)
int bfmmm_host_update_delta_xi(bfmmm_sampler* s) {
  const int K = s->K, P = s->P, M = s->M, D = s->D;
  s->rng.open(HP_DELTA_XI);
  for (int d = 0; d < D; d++)
    for (int k = 0; k < K; k++)
      for (int i = 0; i < M; i++) {
        double p1, p2 = 1;
        if (i == 0) {
          p1 = s->A_xi_(k, 0, d) + ((P * M) * 0.5);
          for (int j = 0; j < P; j++) {
            p2 += 0.5 * s->gamma_xi_(k, j, d, 0) * (s->xi_(k, j, d, 0) * s->xi_(k, j, d, 0));
            for (int m = 1; m < M; m++) {
              double tt = 1;
              for (int nn = 1; nn <= m; nn++) tt *= s->delta_xi_(k, nn, d);
              p2 += 0.5 * s->gamma_xi_(k, j, d, m) * tt * (s->xi_(k, j, d, m) * s->xi_(k, j, d, m));
            }
          }
        } else {
          p1 = s->A_xi_(k, 1, d) + ((P * (M - i)) * 0.5);
          for (int j = 0; j < P; j++)
            for (int m = i; m < M; m++) {
              double tt = 1;
              for (int nn = 0; nn <= m; nn++) if (nn != i) tt *= s->delta_xi_(k, nn, d);
              p2 += 0.5 * s->gamma_xi_(k, j, d, m) * tt * (s->xi_(k, j, d, m) * s->xi_(k, j, d, m));
            }
        }
        s->delta_xi_(k, i, d) = (1 / p2) * s->rng.gamma(p1);
      }
  return tape_check(s);
}

// updateGamma (UpdateGamma.h:17-38)
int bfmmm_host_update_gamma(bfmmm_sampler* s) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  for (int i = 0; i < s->K; i++)
    for (int l = 0; l < s->P; l++) core_update_gamma_row(g, r, i, l);
  return tape_check(s);
}

// updateGammaXi (UpdateGamma.h:48-72)
int bfmmm_host_update_gamma_xi(bfmmm_sampler* s) {
  const int K = s->K, P = s->P, M = s->M, D = s->D;
  s->rng.open(HP_GAMMA_XI);
  const double nug = s->h.nu_1;
  for (int k = 0; k < K; k++)
    for (int i = 0; i < D; i++)
      for (int l = 0; l < P; l++) {
        double ph = 1;
        for (int j = 0; j < M; j++) {
          ph *= s->delta_xi_(k, j, i);
          double scale = 2 / (nug + ph * (s->xi_(k, l, i, j) * s->xi_(k, l, i, j)));
          s->gamma_xi_(k, l, i, j) = scale * s->rng.gamma((nug + 1) / 2);
        }
      }
  return tape_check(s);
}

// updateA (UpdateA.h:58-135; lpdf_a1 :17-23, lpdf_a2 :33-44 in globals_core.cuh)
int bfmmm_host_update_A(bfmmm_sampler* s) {
  if (dev_begin_host_update(s)) return 1;
  GlobalsView g = host_view(s);
  HostRngAdapter r{&s->rng};
  for (int j = 0; j < s->K; j++)
    for (int i = 0; i < 2; i++) core_update_A_one(g, r, j, i);
  return tape_check(s);
}
// updateAXi (UpdateA.h:137-209): order j, i, d
int bfmmm_host_update_A_xi(bfmmm_sampler* s) {
  HostRngAdapter r{&s->rng};
  for (int j = 0; j < s->K; j++)
    for (int i = 0; i < 2; i++)
      for (int d = 0; d < s->D; d++) {
        auto st = r.open(HP_A_XI, (uint64_t)(j * 2 + i) * bf::DMAX + d);
        core_mh_a(s->h, st, s->A_xi_(j, i, d), i == 0, &s->delta_xi_(j, 0, d), s->M, s->K);
      }
  return tape_check(s);
}

// updatePhi (UpdatePhi.h:23-89): blocks (j, m), prior diag(tilde_tau(j,m) * gamma(j,.,m))
int bfmmm_host_update_phi(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta) {
  if (dev_begin_host_update(s)) return 1;
  s->fc_valid = false;
  const int K = s->K, P = s->P, M = s->M;
  s->blk_purpose = HP_PHI; s->blk_index = 0;
  vecd diag(P);
  if (!s->ragged && P >= PREFACTOR_MIN_P) {
    std::vector<PreBlock> blocks;
    for (int j = 0; j < K; j++) {
      double tt = 1;
      for (int m = 0; m < M; m++) {
        tt = (m == 0) ? s->delta_(j, 0) : tt * s->delta_(j, m);
        PreBlock b{s->feat(j, m + 1, 0), nullptr, vecd(P)};
        for (int p = 0; p < P; p++) b.prior_diag[p] = tt * s->gamma_(j, p, m);
        blocks.push_back(std::move(b));
      }
    }
    prefactor_blocks(s, blocks, WtW, beta);
    s->pre_next = 0;
  }
  int rc = 0;
  for (int j = 0; j < K && !rc; j++) {
    double tt = 1;
    for (int m = 0; m < M && !rc; m++) {
      tt = (m == 0) ? s->delta_(j, 0) : tt * s->delta_(j, m);      // tilde_tau cumprod, BFMMM.h:1254-1259
      for (int p = 0; p < P; p++) diag[p] = tt * s->gamma_(j, p, m);
      rc = block_draw(s, j, m + 1, 0, WtW, BtYW, beta, nullptr, diag.data());
    }
  }
  s->pre_next = -1;
  return rc ? rc : tape_check(s);
}
// updateNu (UpdateNu.h:24-74): blocks j, prior tau_j * P (MV: (1/tau_j) I, :195-196)
int bfmmm_host_update_nu(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta) {
  if (dev_begin_host_update(s)) return 1;
  s->fc_valid = false;
  const int K = s->K, P = s->P;
  s->blk_purpose = HP_NU; s->blk_index = 0;
  vecd diag(P);
  const bool pre = !s->ragged && P >= PREFACTOR_MIN_P;
  std::vector<vecd>& priors = s->pre_prior;            // persistent: no 1.3 MB allocations per sweep
  if (!s->identity) {
    if ((int)priors.size() < K) priors.resize(K);
    for (int j = 0; j < K; j++) {
      if (priors[j].size() < (size_t)P * P) priors[j].assign((size_t)P * P, 0.0);
      scale_band(P, s->hbP, s->tau[j], s->Pmat.data(), priors[j].data());
    }
  }
  if (pre) {
    std::vector<PreBlock> blocks;
    for (int j = 0; j < K; j++) {
      if (s->identity) {
        PreBlock b{s->feat(j, 0, 0), nullptr, vecd(P)};
        for (int p = 0; p < P; p++) b.prior_diag[p] = 1 / s->tau[j];
        blocks.push_back(std::move(b));
      } else {
        blocks.push_back(PreBlock{s->feat(j, 0, 0), priors[j].data(), vecd()});
      }
    }
    prefactor_blocks(s, blocks, WtW, beta);
    s->pre_next = 0;
  }
  int rc = 0;
  for (int j = 0; j < K && !rc; j++) {
    if (s->identity) {
      for (int p = 0; p < P; p++) diag[p] = 1 / s->tau[j];
      rc = block_draw(s, j, 0, 0, WtW, BtYW, beta, nullptr, diag.data());
    } else {
      rc = block_draw(s, j, 0, 0, WtW, BtYW, beta, priors[j].data(), nullptr);
    }
  }
  s->pre_next = -1;
  return rc ? rc : tape_check(s);
}
// updateEta (UpdateEta.h:28-94): d outer, j inner; prior tau_eta(j,d) * P (MV: (1/tau_eta) I)
int bfmmm_host_update_eta(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta) {
  if (dev_begin_host_update(s)) return 1;
  s->fc_valid = false;
  const int K = s->K, P = s->P, D = s->D;
  s->blk_purpose = HP_ETA; s->blk_index = 0;
  vecd& prior = s->prior_buf;
  if (prior.size() < (size_t)P * P) prior.assign((size_t)P * P, 0.0);
  vecd diag(P);
  for (int d = 0; d < D; d++)
    for (int j = 0; j < K; j++) {
      if (s->identity) {
        for (int p = 0; p < P; p++) diag[p] = 1 / s->tau_eta_(j, d);
        if (block_draw(s, j, 0, d + 1, WtW, BtYW, beta, nullptr, diag.data())) return 1;
      } else {
        scale_band(P, s->hbP, s->tau_eta_(j, d), s->Pmat.data(), prior.data());
        if (block_draw(s, j, 0, d + 1, WtW, BtYW, beta, prior.data(), nullptr)) return 1;
      }
    }
  return tape_check(s);
}
// updateXiCovariateAdj (UpdateXi.h:26-93): order j, m, d; prior diag(tilde_tau_xi(j,m,d) * gamma_xi_j(.,d,m))
int bfmmm_host_update_xi(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta) {
  if (dev_begin_host_update(s)) return 1;
  s->fc_valid = false;
  const int K = s->K, P = s->P, M = s->M, D = s->D;
  s->blk_purpose = HP_XI; s->blk_index = 0;
  vecd diag(P);
  for (int j = 0; j < K; j++)
    for (int m = 0; m < M; m++)
      for (int d = 0; d < D; d++) {
        double tt = 1;
        for (int mm = 0; mm <= m; mm++) tt *= s->delta_xi_(j, mm, d);
        for (int p = 0; p < P; p++) diag[p] = tt * s->gamma_xi_(j, p, d, m);
        if (block_draw(s, j, m + 1, d + 1, WtW, BtYW, beta, nullptr, diag.data())) return 1;
      }
  return tape_check(s);
}

// updateSigma's draw (UpdateSigma.h:47-53; tempered :98-107; MV :149-151)
int bfmmm_host_update_sigma(bfmmm_sampler* s, double ssr, double beta, int tempered) {
  if (dev_begin_host_update(s)) return 1;
  s->rng.open(HP_SIGMA);
  double a, b1;
  if (tempered) { a = (beta * s->n_points_total) / 2 + s->h.alpha_0; b1 = (beta / 2) * ssr + s->h.beta_0; }
  else { a = s->sum_half_total + s->h.alpha_0; b1 = 0.5 * ssr + s->h.beta_0; }
  double r = (1 / b1) * s->rng.gamma(a);
  s->sigma_sq = 1 / r;
  return tape_check(s);
}

// ================================================================= stored samples
// The reference keeps r_stored_iters slots per parameter, slot = iteration % r_stored_iters, and at
// the end of every batch writes slot 0 and slots thinning*p - 1 (p >= 1) to  dir + Name{q}.txt, then
// moves the last slot to slot 0 (BFMMM.h:1680-1746).  Only the slots that will be written are kept.
static bool slot_is_written(const bfmmm_sampler* s, int slot) {
  if (slot == 0) return true;
  return ((slot + 1) % s->rec.thin) == 0 && (slot + 1) / s->rec.thin < s->rec.r / s->rec.thin;
}
static int record_iteration(bfmmm_sampler* s) {
  auto& R = s->rec;
  const int64_t i = s->iteration;
  const int slot = (int)(i % R.r);
  if (slot_is_written(s, slot)) {
    R.nu[slot] = s->nu; R.Phi[slot] = s->Phi; R.pi[slot] = s->pi; R.delta[slot] = s->delta; R.gamma[slot] = s->gamma;
    R.A[slot] = s->A; R.tau[slot] = s->tau; R.sigma[slot] = s->sigma_sq; R.alpha3[slot] = s->alpha3;
    if (s->D) { R.eta[slot] = s->eta; R.xi[slot] = s->xi; R.tau_eta[slot] = s->tau_eta; R.delta_xi[slot] = s->delta_xi;
                R.gamma_xi[slot] = s->gamma_xi; R.A_xi[slot] = s->A_xi; }
    R.Z[slot].resize((size_t)s->n * s->K); R.chi[slot].resize((size_t)s->n * s->M);
    if (bfmmm_get_state(s->e, R.Z[slot].data(), R.chi[slot].data())) return 1;
  }
  if (((i + 1) % R.r) != 0 || i <= 1) return 0;
  // ---- write the batch
  const int ns = R.r / R.thin;
  const int K = s->K, P = s->P, M = s->M, D = s->D, n = s->n;
  auto slot_of = [&](int p) { return p == 0 ? 0 : R.thin * p - 1; };
  auto gather = [&](const std::vector<vecd>& src, size_t per) {
    vecd out(per * ns);
    for (int p = 0; p < ns; p++) std::copy(src[slot_of(p)].begin(), src[slot_of(p)].begin() + per, out.begin() + per * p);
    return out;
  };
  const std::string q = std::to_string(R.q);
  auto path = [&](const char* name) { return R.dir + name + q + ".txt"; };
  vecd v;
  v = gather(R.nu, (size_t)K * P);    if (bfmmm_save_cube_txt(path("Nu").c_str(), v.data(), K, P, ns)) return 1;
  v = gather(R.chi, (size_t)n * M);   if (bfmmm_save_cube_txt(path("Chi").c_str(), v.data(), n, M, ns)) return 1;
  v = gather(R.pi, (size_t)K);        if (bfmmm_save_mat_txt(path("Pi").c_str(), v.data(), K, ns)) return 1;
  v.assign(ns, 0.0);                  // alpha_31(0) is never assigned in the reference (:1685,1695-1704)
  for (int p = 1; p < ns; p++) v[p] = R.alpha3[slot_of(p)];
  if (bfmmm_save_mat_txt(path("alpha_3").c_str(), v.data(), ns, 1)) return 1;
  v = gather(R.A, (size_t)K * 2);     if (bfmmm_save_cube_txt(path("A").c_str(), v.data(), K, 2, ns)) return 1;
  v = gather(R.delta, (size_t)K * M); if (bfmmm_save_cube_txt(path("Delta").c_str(), v.data(), K, M, ns)) return 1;
  v.assign(ns, 0.0);
  for (int p = 0; p < ns; p++) v[p] = R.sigma[slot_of(p)];
  if (bfmmm_save_mat_txt(path("Sigma").c_str(), v.data(), ns, 1)) return 1;
  v.assign((size_t)ns * K, 0.0);      // tau1 is (draws x K)
  for (int p = 0; p < ns; p++) for (int k = 0; k < K; k++) v[(size_t)k * ns + p] = R.tau[slot_of(p)][k];
  if (bfmmm_save_mat_txt(path("Tau").c_str(), v.data(), ns, K)) return 1;
  v = gather(R.gamma, (size_t)K * P * M); if (bfmmm_save_field_cube_bin(path("Gamma").c_str(), v.data(), ns, 1, K, P, M)) return 1;
  v = gather(R.Phi, (size_t)K * P * M);   if (bfmmm_save_field_cube_bin(path("Phi").c_str(), v.data(), ns, 1, K, P, M)) return 1;
  v = gather(R.Z, (size_t)n * K);     if (bfmmm_save_cube_txt(path("Z").c_str(), v.data(), n, K, ns)) return 1;
  if (D) {                            // covariate-adjusted drivers add these (BFMMM.h:5152-5168)
    v = gather(R.eta, (size_t)P * D * K);  if (bfmmm_save_field_cube_bin(path("Eta").c_str(), v.data(), ns, 1, P, D, K)) return 1;
    // field (draws x K) of P x D x M cubes, column-major over the field: k outer
    auto field_k = [&](const std::vector<vecd>& src) {
      const size_t per = (size_t)P * D * M;
      vecd out(per * ns * K);
      for (int k = 0; k < K; k++)
        for (int p = 0; p < ns; p++)
          std::copy(src[slot_of(p)].begin() + per * k, src[slot_of(p)].begin() + per * (k + 1), out.begin() + per * ((size_t)k * ns + p));
      return out;
    };
    v = field_k(R.xi);        if (bfmmm_save_field_cube_bin(path("Xi").c_str(), v.data(), ns, K, P, D, M)) return 1;
    v = field_k(R.gamma_xi);  if (bfmmm_save_field_cube_bin(path("Gamma_Xi").c_str(), v.data(), ns, K, P, D, M)) return 1;
    v = gather(R.delta_xi, (size_t)K * M * D); if (bfmmm_save_field_cube_bin(path("Delta_Xi").c_str(), v.data(), ns, 1, K, M, D)) return 1;
    v = gather(R.A_xi, (size_t)K * 2 * D);     if (bfmmm_save_field_cube_bin(path("A_Xi").c_str(), v.data(), ns, 1, K, 2, D)) return 1;
    v = gather(R.tau_eta, (size_t)K * D);      if (bfmmm_save_cube_txt(path("Tau_Eta").c_str(), v.data(), K, D, ns)) return 1;
  }
  // slot 0 <- last slot (:1733-1745)
  R.q++;
  return 0;
}

}  // extern "C"
namespace { int record_iteration_fwd(bfmmm_sampler* s) { return record_iteration(s); } }
extern "C" {

// ================================================================= driver loops
static int sampler_step_impl(bfmmm_sampler* s, int sweep, double beta);
int bfmmm_sampler_step(bfmmm_sampler* s, int sweep, double beta) {
  if (!s) return sfail("null sampler");
  if (!s->e) return sfail("bfmmm_sampler_step: detached sampler has no engine");
  const double t0 = now_s(), w0 = s->t_wait, p0 = s->t_push;
  int rc = s->dev.on ? sampler_step_device(s, sweep, beta) : sampler_step_impl(s, sweep, beta);
  s->t_host += (now_s() - t0) - (s->t_wait - w0) - (s->t_push - p0);
  if (s->rng.tape_underrun) { s->rng.tape_underrun = false; if (!rc) rc = sfail("bfmmm_sampler_step: the tape of injected draws ran out during the sweep"); }
  return rc;
}
// seconds spent so far in: host-side draws | waiting on the device | pushing globals
int bfmmm_sampler_profile(bfmmm_sampler* s, double* out3) {
  if (!s) return sfail("null sampler");
  out3[0] = s->t_host; out3[1] = s->t_wait; out3[2] = s->t_push;
  return 0;
}
static int sampler_step_impl(bfmmm_sampler* s, int sweep, double beta) {
  bfmmm_engine* e = s->e;
  s->rng.iteration = (uint64_t)s->tick;
  const bool do_z = (sweep == BFMMM_SWEEP_NU_Z || sweep == BFMMM_SWEEP_FULL);
  const bool do_phi = (sweep == BFMMM_SWEEP_THETA || sweep == BFMMM_SWEEP_FULL);
  const bool do_nu = do_z;
  const bool do_chi = do_phi;
  // updateSigmaTempered's shape a = sum_i beta n_i / 2 (real division, UpdateSigma.h:98-107) is used on EVERY rung of a
  // tempered transition, the beta = 1 rungs included (BFMMM.h:1556-1651); updateSigma's sum_i floor(n_i / 2) otherwise
  const bool tempered = s->in_tt || beta != 1.0;
  if (bfmmm_seed(e, s->rng.key, (uint64_t)s->tick)) return 1;
  if (push_globals(s)) return 1;
  // Z step in two kernels (common basis): the proposal of sweep t + 1 needs only Z, pi and alpha_3 of sweep t, so it is
  // queued on a side stream as soon as pi and alpha_3 are known and runs beside the statistics kernel and the host's
  // Gaussian block draws.  pi and alpha_3 draw from their own Philox streams, so taking them first changes nothing
  // unless the draws come from a tape (then the reference's order is kept and nothing runs ahead).
  // It pays when the gap in which the device would idle (the host's block draws plus ~25 us of read-back, push and launch
  // latency) is about as long as the proposal kernel; with quick block draws (identity basis) the kernel would only be
  // pushed behind the SSR pass by the stream
  // priorities and the Z step would wait for it, so it then stays in front of its accept kernel.  Both times are
  // measured: the first sweep's proposal runs alone on the engine's stream (timed by events), the block draws by the clock.
  static const bool z_ahead_on = !std::getenv("BFMMM_NO_Z_AHEAD"), z_ahead_always = std::getenv("BFMMM_Z_AHEAD_ALWAYS") != nullptr;
  {
    // the shortest of the timings seen: the first launch of a kernel also waits for its module to load
    const double t = bfmmm_z_propose_us(e);
    if (t >= 0 && (s->z_k1_us < 0 || t < s->z_k1_us)) s->z_k1_us = t;
  }
  const bool worth = z_ahead_always || (s->z_k1_us >= 0 && s->z_blocks_us >= 0 && s->z_blocks_us + 25.0 >= 0.7 * s->z_k1_us);
  const bool z_ahead = do_z && !s->rng.use_tape && z_ahead_on && worth && bfmmm_z_ahead_supported(e);
  const bool z_early = z_ahead && !s->allreduce;           // one shard: sum_i log Z_ik needs no exchange
  if (do_z) {                                              // updateZ_PM -> updatePi_PM -> updateAlpha3
    if (bfmmm_update_z_async(e, s->pi.data(), s->alpha3, s->h.a_Z_PM, beta)) return 1;
    if (z_early && bfmmm_slz_read_begin(e)) return 1;
  }
  // the sufficient statistics depend only on (Z, chi, X): one pass feeds Phi, nu, eta and xi
  if (bfmmm_suffstats_async(e)) return 1;
  // While the device runs the Z and statistics kernels: the standard normals of the Phi and nu block draws
  // (they do not depend on the statistics), from the streams and in the order the draws would use.
  s->zpre.clear(); s->zpre_pos = 0;
  if (!s->rng.use_tape) {
    HostRngAdapter r{&s->rng};
    for (int upd = 0; upd < 2; upd++) {
      if (upd == 0 ? !do_phi : !do_nu) continue;
      const int nblk = upd == 0 ? s->K * s->M : s->K;
      for (int t = 0; t < nblk; t++)
        for (int p = 0; p < s->P; p++) { auto st = r.open(upd == 0 ? HP_PHI : HP_NU, ((uint64_t)t << 12) + (uint64_t)p); s->zpre.push_back(st.normal()); }
    }
  }
  if (z_early) {                                           // while the statistics kernel runs
    double slz[9];
    { struct T { bfmmm_sampler* s; double t0; ~T() { s->t_wait += now_s() - t0; } } timer{s, now_s()};
      if (bfmmm_slz_read_wait(e, slz)) return 1; }
    if (bfmmm_host_update_pi(s, slz)) return 1;
    if (bfmmm_host_update_alpha3(s, slz)) return 1;
    static const bool beside = std::getenv("BFMMM_Z_AHEAD_BESIDE") != nullptr;   // experiment: run beside the statistics kernel
    if (bfmmm_z_propose_async(e, s->pi.data(), s->alpha3, s->h.a_Z_PM, (uint64_t)(s->tick + 1), !beside)) return 1;
  }
  if (reduce_and_read(s)) return 1;
  if (s->ll_pending) {                                     // previous sweep's post-chi SSR arrived with this exchange
    s->last_ssr = st_ssr_after(s);
    set_loglik(s, s->last_ssr, s->ll_sigma);
    s->ll_pending = false;
  }
  if (s->ragged) s->Hb = st_hb(s);
  if (do_z) s->last_accept = (int64_t)std::llround(st_acc(s));
  // The Gaussian blocks that change the mean (Phi, nu) are drawn first so that the SSR pass can start;
  // the prior updates that do not feed this sweep's device passes (pi, alpha_3, delta, A, gamma, tau)
  // then run on the host WHILE the device streams the SSR pass.  Each update draws from its own
  // Philox stream and reads exactly what it reads in the reference's order (Phi uses the previous
  // delta/gamma, nu the previous tau; delta, A, gamma see the new Phi; tau the new nu), so the chain is
  // the one of BFMMM.h:1500-1554 -- only wall-clock placement differs.
  if (z_ahead && !z_early) {                               // several shards: the sums have just been exchanged
    if (bfmmm_host_update_pi(s, st_slz(s))) return 1;
    if (bfmmm_host_update_alpha3(s, st_slz(s))) return 1;
    if (bfmmm_z_propose_async(e, s->pi.data(), s->alpha3, s->h.a_Z_PM, (uint64_t)(s->tick + 1), false)) return 1;
  }
  s->zpre_on = !s->rng.use_tape;
  const double t_blocks0 = now_s();
  int rc_blocks = (do_phi && bfmmm_host_update_phi(s, st_wtw(s), st_btyw(s), beta)) ||   // updatePhi
                  (do_nu && bfmmm_host_update_nu(s, st_wtw(s), st_btyw(s), beta));        // updateNu
  s->zpre_on = false;
  if (do_phi && do_nu) {
    const double us = 1e6 * (now_s() - t_blocks0);
    s->z_blocks_us = s->z_blocks_us < 0 ? us : 0.8 * s->z_blocks_us + 0.2 * us;
  }
  if (rc_blocks) return 1;
  if (push_globals(s)) return 1;
  // Fused path (no injected draws, no covariates): sigma^2 is drawn ON THE DEVICE right behind the SSR pass
  // -- same Philox stream (key, iteration, HP_SIGMA), same Marsaglia-Tsang sampler as the host draw -- and the
  // chi kernel is queued immediately behind it, so the device does not idle for a host round trip between
  // the two passes.  The host takes (SSR, sigma^2) from mapped memory when it needs them.  Where the engine can, the
  // draw (and the exchange of the SSR slot over the shards) is the tail of the SSR pass itself: one launch, not three.
  const bool fused_sigma = do_chi && !s->D && !s->rng.use_tape && !std::getenv("BFMMM_NO_FUSED_SIGMA");
  const double a_sh = tempered ? (beta * s->n_points_total) / 2 + s->h.alpha_0 : s->sum_half_total + s->h.alpha_0;
  int tail_done = 0;
  if (fused_sigma && bfmmm_ssr_sigma_async(e, s->allreduce ? 1 : 0, a_sh, tempered ? beta / 2 : 0.5, s->h.beta_0, s->rng.key,
                                           s->rng.iteration, HP_SIGMA, &tail_done)) return 1;
  if (!tail_done && bfmmm_ssr_async(e)) return 1;          // updateSigma's data pass, new globals
  if (fused_sigma) {
    if (!tail_done) {
      double* dev = nullptr; int64_t len = 0;
      if (bfmmm_stats_buffer_dev(e, &dev, &len)) return 1;
      if (s->allreduce && s->allreduce(s->allreduce_ctx, dev + s->K + 1, 1, bfmmm_stream(e))) return sfail("all-reduce hook failed");
      if (bfmmm_sigma_draw_async(e, a_sh, tempered ? beta / 2 : 0.5, s->h.beta_0, s->rng.key, s->rng.iteration, HP_SIGMA)) return 1;
    }
    if (bfmmm_update_chi_async(e, beta)) return 1;
  }
  if (do_z && !z_ahead) {                                  // updatePi_PM -> updateAlpha3
    if (bfmmm_host_update_pi(s, st_slz(s))) return 1;
    if (bfmmm_host_update_alpha3(s, st_slz(s))) return 1;
  }
  if (do_phi) {                                            // updateDelta, updateA, updateGamma
    if (bfmmm_host_update_delta(s)) return 1;
    if (bfmmm_host_update_A(s)) return 1;
    if (bfmmm_host_update_gamma(s)) return 1;
  }
  if (bfmmm_host_update_tau(s)) return 1;                  // updateTau (all three loops call it)
  double ssr_ll;
  if (fused_sigma) {
    struct T { bfmmm_sampler* s; double t0; ~T() { s->t_wait += now_s() - t0; } } timer{s, now_s()};
    double ssr_now = 0, sig_now = 0;
    if (bfmmm_sigma_wait(e, &ssr_now, &sig_now)) return 1;
    s->sigma_sq = sig_now;
    if ((int64_t)s->stats.size() > s->K + 1) s->stats[s->K + 1] = ssr_now;
    ssr_ll = ssr_now;
  } else {
    if (reduce_and_read(s, /*only_ssr=*/true)) return 1;
    if (bfmmm_host_update_sigma(s, st_ssr(s), beta, tempered)) return 1;
    ssr_ll = st_ssr(s);
  }
  bool defer = false;
  if (do_chi) {                                            // updateChi (+ the SSR calcLikelihood needs)
    if (!fused_sigma) {
      if (push_globals(s)) return 1;
      if (bfmmm_update_chi_async(e, beta)) return 1;
    }
    if (!s->D) {
      if (s->in_tt) {                                      // a tempered transition needs every slot's SSR now
        s->ll_pending = true; s->ll_sigma = s->sigma_sq;
        if (flush_loglik(s)) return 1;
        ssr_ll = s->last_ssr;
      } else {
        defer = true;                                      // read with the next sweep's first exchange
      }
    }
  }
  if (s->D) {
    // covariate-adjusted loops (BFMMM.h:3976-4000): after chi come updateEta, updateTauEta and the
    // xi block with its shrinkage priors -- chi has changed, so the statistics are taken again
    if (bfmmm_suffstats_async(e)) return 1;
    if (reduce_and_read(s)) return 1;
    if (s->ragged) s->Hb = st_hb(s);
    if (do_nu && bfmmm_host_update_eta(s, st_wtw(s), st_btyw(s), beta)) return 1;
    if (bfmmm_host_update_tau_eta(s)) return 1;
    if (do_phi) {
      if (bfmmm_host_update_xi(s, st_wtw(s), st_btyw(s), beta)) return 1;
      if (bfmmm_host_update_delta_xi(s)) return 1;
      if (bfmmm_host_update_A_xi(s)) return 1;
      if (bfmmm_host_update_gamma_xi(s)) return 1;
    }
    if (push_globals(s)) return 1;
    if (bfmmm_ssr_async(e)) return 1;
    if (reduce_and_read(s, /*only_ssr=*/true)) return 1;
    ssr_ll = st_ssr(s);
  }
  if (defer) {
    s->ll_pending = true; s->ll_sigma = s->sigma_sq;
  } else {
    s->last_ssr = ssr_ll;
    set_loglik(s, ssr_ll, s->sigma_sq);
  }
  s->tick++;
  if (!s->in_tt) {
    if (s->rec.on && record_iteration(s)) return 1;
    s->iteration++;
  }
  return 0;
}

// One tempered transition (BFMMM.h:1556-1651; ladder :1452-1460; acceptance
// CalculateTTAcceptance.h:22-97): 2 N_t tempered full sweeps up and down the ladder, then
//   log A = sum_i (beta_{i+1} - beta_i) [ g(s_i) - g(s_{m-i}) ],  g(s) = -(N/2) log sigma_s^2 - SSR_s / (2 sigma_s^2),
// accept iff log u < log A, otherwise every parameter (and Z, chi on the device) is restored.
int bfmmm_sampler_tempered_transition(bfmmm_sampler* s, int N_t, double beta_N_t, double* logA_out, int* accepted) {
  if (!s || !s->e) return sfail("null sampler");
  if (N_t < 1) return sfail("bfmmm_sampler_tempered_transition: N_t must be >= 1");
  std::vector<double> ladder(N_t, 1.0);
  ladder[N_t - 1] = beta_N_t;
  const double geom = std::pow(beta_N_t, 1.0 / N_t);
  for (int i = 1; i < N_t; i++) ladder[i] = ladder[i - 1] * geom;      // as written at BFMMM.h:1453-1460
  const int m = 2 * N_t;
  if (dev_pull(s)) return 1;
  if (flush_loglik(s)) return 1;
  // slot 0: the current state.  Its SSR: one data pass with the current globals.
  if (push_globals(s)) return 1;
  if (bfmmm_ssr_async(s->e)) return 1;
  if (reduce_and_read(s, /*only_ssr=*/true)) return 1;
  s->tt_ssr.assign(m + 1, 0.0); s->tt_sigma.assign(m + 1, 0.0);
  s->tt_ssr[0] = st_ssr(s); s->tt_sigma[0] = s->sigma_sq;
  const ChainParams saved = save_params(s);       // the chain's globals only (not the recorder's kept draws or scratch)
  if (bfmmm_state_snapshot(s->e)) return 1;
  int temp_ind = 0;
  s->in_tt = true;
  for (int l = 1; l <= m; l++) {
    if (bfmmm_sampler_step(s, BFMMM_SWEEP_FULL, ladder[temp_ind])) {
      // a failed rung leaves nothing half-advanced: globals and (Z, chi) go back to the pre-transition state
      s->in_tt = false;
      restore_params(s, saved);
      s->dev.dev_stale = true; s->dev.host_stale = false;
      bfmmm_state_restore(s->e);
      return 1;
    }
    s->tt_ssr[l] = s->last_ssr; s->tt_sigma[l] = s->sigma_sq;
    if (l < N_t) temp_ind++;
    if (l > N_t) temp_ind--;
  }
  s->in_tt = false;
  double logA = 0;
  auto g = [&](double beta, int slot) {
    return -(beta / 2) * s->n_points_total * std::log(s->tt_sigma[slot]) - (beta / (2 * s->tt_sigma[slot])) * s->tt_ssr[slot];
  };
  for (int i = 0; i + 1 < N_t; i++) {
    logA += g(ladder[i + 1], i) - g(ladder[i], i);
    logA += -g(ladder[i + 1], m - i) + g(ladder[i], m - i);
  }
  s->rng.iteration = (uint64_t)s->tick;
  s->rng.open(HP_TT);
  s->tick++;
  const double logu = std::log(s->rng.uniform());
  const bool ok = logu < logA;
  s->tt_total++;
  if (ok) s->tt_accepts++;
  else {
    // restore the pre-transition parameters; the bookkeeping (tick, counters, traces, random streams) moves on
    if (dev_pull(s)) return 1;
    restore_params(s, saved);
    s->dev.dev_stale = true;
    if (bfmmm_state_restore(s->e)) return 1;
  }
  if (s->rec.on && record_iteration(s)) return 1;
  s->iteration++;                                  // a transition is one outer iteration (BFMMM.h:1500)
  if (logA_out) *logA_out = logA;
  if (accepted) *accepted = ok ? 1 : 0;
  return 0;
}

// BFMMM_MTT_warm_start's iteration schedule (BFMMM.h:1500-1556): a plain sweep unless i is a positive
// multiple of n_temp_trans, in which case the sweep is replaced by a tempered transition.
int bfmmm_sampler_run_mtt(bfmmm_sampler* s, int n_iter, int n_temp_trans, int N_t, double beta_N_t) {
  if (!s) return sfail("null sampler");
  if (n_temp_trans <= 0) n_temp_trans = n_iter + 1;          // UserFunctions.cpp:1353-1359
  for (int it = 0; it < n_iter; it++) {
    const int64_t i = s->iteration;
    if ((i % n_temp_trans) != 0 || i == 0) {
      if (bfmmm_sampler_step(s, BFMMM_SWEEP_FULL, 1.0)) return 1;
    } else {
      if (bfmmm_sampler_tempered_transition(s, N_t, beta_N_t, nullptr, nullptr)) return 1;
    }
  }
  return 0;
}
// start writing the reference's stored-sample files: a batch every r_stored_iters iterations,
// keeping every thinning_num-th draw
int bfmmm_sampler_record(bfmmm_sampler* s, const char* directory, int r_stored_iters, int thinning_num) {
  if (!s || !s->e) return sfail("null sampler");
  if (!directory || r_stored_iters < 2 || thinning_num < 1 || r_stored_iters / thinning_num < 1)
    return sfail("bfmmm_sampler_record: need a directory, r_stored_iters >= 2 and 1 <= thinning_num <= r_stored_iters");
  auto& R = s->rec;
  R.on = true; R.dir = directory; R.r = r_stored_iters; R.thin = thinning_num; R.q = 0;
  for (auto* v : {&R.nu, &R.Phi, &R.pi, &R.delta, &R.gamma, &R.A, &R.tau, &R.Z, &R.chi, &R.eta, &R.xi, &R.tau_eta,
                  &R.delta_xi, &R.gamma_xi, &R.A_xi})
    v->assign(r_stored_iters, vecd());
  R.sigma.assign(r_stored_iters, 0.0); R.alpha3.assign(r_stored_iters, 0.0);
  return 0;
}
int bfmmm_sampler_batches_written(bfmmm_sampler* s) { return s ? s->rec.q : -1; }

int bfmmm_sampler_tt_trace(bfmmm_sampler* s, double* ssr, double* sigma, int n) {
  if (!s) return sfail("null sampler");
  for (int i = 0; i < n && i < (int)s->tt_ssr.size(); i++) { ssr[i] = s->tt_ssr[i]; sigma[i] = s->tt_sigma[i]; }
  return (int)s->tt_ssr.size() == 0 ? sfail("no tempered transition has run") : 0;
}

int bfmmm_sampler_run(bfmmm_sampler* s, int sweep, int n_iter) {
  for (int i = 0; i < n_iter; i++)
    if (bfmmm_sampler_step(s, sweep, 1.0)) return 1;
  return 0;
}

}  // extern "C"
