// engine.cu -- the C ABI of include/bfmmm.h: device memory, create-time projection, launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bfmmm.h"
#include "common.cuh"
#include "engine_internal.h"

namespace bf { std::atomic<unsigned long long> g_launch_count{0}; }

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }
}
namespace bf { int set_error(const char* msg) { g_err = msg; return 1; } }
namespace {
#define CU(x)                                                                          \
  do {                                                                                 \
    cudaError_t _e = (x);                                                              \
    if (_e != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(_e)); \
  } while (0)

constexpr int N_STAGE = 4;
}  // namespace

struct bfmmm_engine {
  int model = 0, n = 0, ld = 0, K = 0, P = 0, M = 0, D = 0, q = 0, QS = 0, device = 0;
  int hbL = 0;  // lower bandwidth of the whitening factor L (B-spline Gram: banded), used by the whitening loops
  int P4 = 0;   // P rounded up to a multiple of 4: allocated rows of Ct / glob (rows >= Pc stay zero)
  int Pc = 0;   // rows of the projected cache = rank of the basis Gram (= P unless the basis is rank deficient)
  int64_t T = 0;
  bool common = true, identity = false;
  int64_t global_offset = 0;
  double n_points = 0, sum_half = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  // device
  double *Ct = nullptr, *rss = nullptr, *Z = nullptr, *chi = nullptr, *X = nullptr, *glob = nullptr;
  double *Hh = nullptr, *Gl = nullptr, *rs_partials = nullptr;   // ragged grids: B_i'y_i, band of G_i
  int bw = 0, npairs = 0;
  bool ragged = false;
  double *snapZ = nullptr, *snapChi = nullptr;   // device copy of (Z, chi) for tempered transitions
  // per-function moments (r[0..M-1], rss + |c~ - mu~|^2) left by the SSR pass for the chi step that follows it
  // (moments_kernels.cu); valid until Z, the globals or the data change
  double* mom = nullptr;
  bool mom_valid = false, mom_enabled = false;
  // Z step in two halves (z_propose_kernel / z_kernel<PRE>): the proposal of the next sweep may be made ahead of time on
  // a side stream (bfmmm_z_propose_async); it is used when the step is asked for with exactly the parameters it was made with
  double* zprop = nullptr;
  bool z_split = false, prop_valid = false;
  cudaStream_t side = nullptr;
  int prio_side = 0;
  double* h_slz = nullptr;        // mapped page-locked copy of [sum log Z (K) | accepts] read right behind the Z step
  double* h_slz_dev = nullptr;
  cudaEvent_t ev_slz = nullptr;
  cudaEvent_t ev_zdone = nullptr, ev_prop = nullptr, ev_queued = nullptr;
  cudaEvent_t ev_k1a = nullptr, ev_k1b = nullptr;      // timing events around a proposal kernel on the engine's own stream
  bool k1_timed = false;
  long long n_prop_ahead = 0, n_prop_own = 0;     // Z steps that used an ahead-of-time proposal / made their own
  double prop_pi[8] = {0}, prop_alpha3 = 0, prop_a = 0;
  uint64_t prop_key = 0, prop_iter = 0;
  double *ni = nullptr;                          // ragged grids: points per function (marginal log-likelihood)
  double *cpo_m = nullptr, *cpo_s = nullptr, *logl = nullptr;   // CPO accumulators, per-function marginal log-likelihood
  int64_t cpo_count = 0;
  double *rbZ = nullptr, *rbChi = nullptr;       // device staging of an overlapped read-back (bfmmm_get_state_begin)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
  bool rb_pending = false;
  double *draws = nullptr, *stats = nullptr, *partials = nullptr, *st_partials = nullptr, *acc_dbg = nullptr;
  unsigned int* ticket = nullptr;
  int64_t stats_len = 0;
  bf::StatsTmaMaps tma;            // TMA descriptors of the statistics kernel (valid = 0: cp.async path)
  int pass_blocks = 0, st_blocks = 0;
  // host
  std::vector<double> B, G, L;       // basis T x P row-major, Gram P x P col-major, whitening matrix P x Pc col-major
                                     // (L = chol_lower(G), or V_r Lambda_r^{1/2} for a rank-deficient Gram): G = L L'
  double* h_stage[N_STAGE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_stage[N_STAGE] = {nullptr, nullptr, nullptr, nullptr};
  int stage_next = 0;
  double* h_stats = nullptr;
  double* h_stats_dev = nullptr;   // device alias of h_stats (mapped page-locked memory)
  bool mirror_valid = false;       // h_stats holds the whole reduced buffer (stored by the statistics pass's epilogue)
  bool stats_exchanged = false;    // ... and that epilogue already summed it over the shards
  unsigned int* st_ticket = nullptr;
  bool xchg_on = false;            // peer-memory exchange fused into the statistics pass's epilogue (bfmmm_engine_set_exchange)
  bf::P2PPeers xchg_peers; int xchg_rank = 0, xchg_world = 1, xchg_cap = 0;
  unsigned long long* xchg_seq = nullptr;
  double *sigma_dev = nullptr, *h_sig = nullptr, *h_sig_dev = nullptr;   // device-drawn sigma^2; mapped {SSR, sigma^2, seq}
  double sig_seq = 0;
  bool sigma_armed = false;        // the next chi kernel reads sigma^2 from sigma_dev
  double sigma_sq = 1.0;
  uint64_t key = 0x9E3779B97F4A7C15ull, iteration = 0;
  // offsets into stats
  int off_slz() const { return 0; }
  int off_acc() const { return K; }
  int off_ssr() const { return K + 1; }
  int off_ssr_after() const { return K + 2; }
  int off_wtw() const { return K + 3; }
  int off_ctw() const { return K + 3 + q * q; }
  int off_hb() const { return K + 3 + q * q + P * q; }
};

namespace {

bool chol_lower(int n, const std::vector<double>& A, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) s -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
    if (!(s > 0)) return false;
    double d = std::sqrt(s);
    L[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double t = A[(size_t)j * n + i];
      for (int k = 0; k < j; k++) t -= L[(size_t)k * n + i] * L[(size_t)k * n + j];
      L[(size_t)j * n + i] = t / d;
    }
  }
  return true;
}

void free_all(bfmmm_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaFree(e->Ct); cudaFree(e->rss); cudaFree(e->Z); cudaFree(e->chi); cudaFree(e->X); cudaFree(e->glob);
  cudaFree(e->draws); cudaFree(e->stats); cudaFree(e->partials); cudaFree(e->st_partials); cudaFree(e->ticket); cudaFree(e->st_ticket);
  cudaFree(e->acc_dbg); cudaFree(e->snapZ); cudaFree(e->snapChi); cudaFree(e->mom); cudaFree(e->zprop);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->ev_zdone) cudaEventDestroy(e->ev_zdone);
  if (e->ev_prop) cudaEventDestroy(e->ev_prop);
  if (e->ev_queued) cudaEventDestroy(e->ev_queued);
  if (e->ev_k1a) cudaEventDestroy(e->ev_k1a);
  if (e->ev_k1b) cudaEventDestroy(e->ev_k1b);
  if (e->ev_slz) cudaEventDestroy(e->ev_slz);
  if (e->h_slz) cudaFreeHost(e->h_slz);
  cudaFree(e->Hh); cudaFree(e->Gl); cudaFree(e->rs_partials);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  if (e->ev_snap) cudaEventDestroy(e->ev_snap);
  if (e->ev_copied) cudaEventDestroy(e->ev_copied);
  cudaFree(e->rbZ); cudaFree(e->rbChi);
  cudaFree(e->ni); cudaFree(e->cpo_m); cudaFree(e->cpo_s); cudaFree(e->logl);
  for (int i = 0; i < N_STAGE; i++) {
    if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
    if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]);
  }
  if (e->h_stats) cudaFreeHost(e->h_stats);
  if (e->h_sig) cudaFreeHost(e->h_sig);
  cudaFree(e->sigma_dev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

// host (n x cols, column-major, leading dimension n) -> device [cols][ld]
int upload_cols(bfmmm_engine* e, double* dst, const double* src, int cols) {
  if (cols == 0) return 0;
  CU(cudaMemcpy2DAsync(dst, (size_t)e->ld * 8, src, (size_t)e->n * 8, (size_t)e->n * 8, cols, cudaMemcpyHostToDevice, e->stream));
  return 0;
}
int download_cols(bfmmm_engine* e, double* dst, const double* src, int cols) {
  if (cols == 0) return 0;
  CU(cudaMemcpy2DAsync(dst, (size_t)e->n * 8, src, (size_t)e->ld * 8, (size_t)e->n * 8, cols, cudaMemcpyDeviceToHost, e->stream));
  return 0;
}

int build_basis(bfmmm_engine* e, const bfmmm_config* c) {
  const int P = e->P;
  const int64_t T = e->T;
  e->B.assign((size_t)T * P, 0.0);
  if (c->B) {
    std::copy(c->B, c->B + (size_t)T * P, e->B.begin());
  } else {
    if (!c->t || (c->n_internal > 0 && !c->internal_knots)) return fail("bfmmm_create: neither B nor a spline description (t, knots) was given");
    if (c->n_internal + c->degree + 1 != P) return fail("bfmmm_create: P != n_internal + degree + 1");
    int nk = c->n_internal + 2 * (c->degree + 1);
    std::vector<double> kn(nk);
    for (int i = 0; i <= c->degree; i++) { kn[i] = c->boundary[0]; kn[nk - 1 - i] = c->boundary[1]; }
    for (int i = 0; i < c->n_internal; i++) kn[c->degree + 1 + i] = c->internal_knots[i];
    double *d_t = nullptr, *d_kn = nullptr, *d_B = nullptr;
    CU(cudaMalloc(&d_t, T * 8)); CU(cudaMalloc(&d_kn, nk * 8)); CU(cudaMalloc(&d_B, (size_t)T * P * 8));
    CU(cudaMemcpyAsync(d_t, c->t, T * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_kn, kn.data(), nk * 8, cudaMemcpyHostToDevice, e->stream));
    if (bf::launch_bspline(d_t, T, d_kn, nk, c->degree, P, d_B, e->stream)) return fail("bspline kernel launch failed");
    CU(cudaMemcpyAsync(e->B.data(), d_B, (size_t)T * P * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_t); cudaFree(d_kn); cudaFree(d_B);
  }
  // Gram and its Cholesky factor
  e->G.assign((size_t)P * P, 0.0);
  for (int64_t t = 0; t < T; t++) {
    const double* b = &e->B[(size_t)t * P];
    for (int cc = 0; cc < P; cc++) {
      if (b[cc] == 0.0) continue;
      for (int r = 0; r < P; r++) e->G[(size_t)cc * P + r] += b[cc] * b[r];
    }
  }
  e->Pc = P;
  if (!chol_lower(P, e->G, e->L)) {
    // rank-deficient Gram (a basis function without support on the grid, or T < P): G = V Lambda V',
    // whitening matrix L := V_r Lambda_r^{1/2} (P x r) still satisfies G = L L' and B theta = Q (L' theta)
    // with Q = B V_r Lambda_r^{-1/2}, so every identity the kernels rely on holds with r cache rows.
    std::vector<double> a(e->G), V((size_t)P * P, 0.0);
    for (int i = 0; i < P; i++) V[(size_t)i * P + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
      double off = 0;
      for (int p = 0; p < P; p++) for (int q2 = p + 1; q2 < P; q2++) off += a[(size_t)q2 * P + p] * a[(size_t)q2 * P + p];
      if (off < 1e-300) break;
      for (int p = 0; p < P; p++)
        for (int q2 = p + 1; q2 < P; q2++) {
          double apq = a[(size_t)q2 * P + p];
          if (apq == 0.0) continue;
          double theta = (a[(size_t)q2 * P + q2] - a[(size_t)p * P + p]) / (2 * apq);
          double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
          double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
          for (int k = 0; k < P; k++) { double x = a[(size_t)p * P + k], y = a[(size_t)q2 * P + k]; a[(size_t)p * P + k] = cs * x - sn * y; a[(size_t)q2 * P + k] = sn * x + cs * y; }
          for (int k = 0; k < P; k++) { double x = a[(size_t)k * P + p], y = a[(size_t)k * P + q2]; a[(size_t)k * P + p] = cs * x - sn * y; a[(size_t)k * P + q2] = sn * x + cs * y; }
          for (int k = 0; k < P; k++) { double x = V[(size_t)p * P + k], y = V[(size_t)q2 * P + k]; V[(size_t)p * P + k] = cs * x - sn * y; V[(size_t)q2 * P + k] = sn * x + cs * y; }
        }
    }
    double lmax = 0;
    for (int i = 0; i < P; i++) lmax = std::max(lmax, a[(size_t)i * P + i]);
    if (!(lmax > 0)) return fail("bfmmm_create: the basis is identically zero on the grid");
    e->L.assign((size_t)P * P, 0.0);
    int r = 0;
    for (int j = 0; j < P; j++) {
      double lam = a[(size_t)j * P + j];
      if (lam <= 1e-10 * lmax) continue;
      for (int i = 0; i < P; i++) e->L[(size_t)r * P + i] = V[(size_t)j * P + i] * std::sqrt(lam);
      r++;
    }
    e->Pc = r;
    e->L.resize((size_t)P * r);
  }
  return 0;
}

int project_common(bfmmm_engine* e, const bfmmm_config* c) {
  const int P = e->P;
  const int64_t T = e->T;
  // Q (T x Pc) with orthonormal columns and B = Q L': full rank: row t of Q solves L q = b_t;
  // rank-deficient: Q = B L (L'L)^{-1} with L'L = Lambda_r diagonal
  const int Pc = e->Pc;
  std::vector<double> Q((size_t)T * Pc);
  if (Pc == P) {
    for (int64_t t = 0; t < T; t++) {
      const double* b = &e->B[(size_t)t * P];
      double* qv = &Q[(size_t)t * P];
      for (int i = 0; i < P; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= e->L[(size_t)k * P + i] * qv[k];
        qv[i] = s / e->L[(size_t)i * P + i];
      }
    }
  } else {
    std::vector<double> lam(Pc, 0.0);
    for (int j = 0; j < Pc; j++) for (int i = 0; i < P; i++) lam[j] += e->L[(size_t)j * P + i] * e->L[(size_t)j * P + i];
    for (int64_t t = 0; t < T; t++) {
      const double* b = &e->B[(size_t)t * P];
      for (int j = 0; j < Pc; j++) {
        double s = 0;
        for (int i = 0; i < P; i++) s += b[i] * e->L[(size_t)j * P + i];
        Q[(size_t)t * Pc + j] = s / lam[j];
      }
    }
  }
  double* d_Q = nullptr;
  CU(cudaMalloc(&d_Q, (size_t)T * Pc * 8));
  CU(cudaMemcpyAsync(d_Q, Q.data(), (size_t)T * Pc * 8, cudaMemcpyHostToDevice, e->stream));
  // stream the observations through a bounded device buffer
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(e->n, (int64_t)(1ull << 30) / (T * 8)));
  double* d_Y = nullptr;
  CU(cudaMalloc(&d_Y, (size_t)chunk * T * 8));
  for (int64_t i0 = 0; i0 < e->n; i0 += chunk) {
    int64_t m = std::min<int64_t>(chunk, e->n - i0);
    CU(cudaMemcpyAsync(d_Y, c->y + i0 * T, (size_t)m * T * 8, cudaMemcpyHostToDevice, e->stream));
    bf::ProjectArgs pa;
    pa.n = (int)m; pa.ld = e->ld; pa.P = Pc; pa.T = T; pa.i_begin = i0; pa.Y = d_Y; pa.Q = d_Q; pa.Ct = e->Ct; pa.rss = e->rss;
    if (bf::launch_project(pa, e->stream)) return fail("project kernel launch failed");
    CU(cudaStreamSynchronize(e->stream));
  }
  cudaFree(d_Y); cudaFree(d_Q);
  return 0;
}

// ragged grids: per-function banded Gram, least-squares coefficients, B_i'y_i, orthogonal residual
int project_ragged(bfmmm_engine* e, const bfmmm_config* c) {
  const int P = e->P, n = e->n;
  if (!c->off) return fail("bfmmm_create: ragged grids need the offsets `off`");
  const int64_t N = c->off[n] - c->off[0];
  if (N <= 0) return fail("bfmmm_create: empty observations");
  double sum_half = 0;
  for (int i = 0; i < n; i++) {
    int64_t ni = c->off[i + 1] - c->off[i];
    if (ni < 0) return fail("bfmmm_create: offsets must be non-decreasing");
    sum_half += (double)(ni / 2);                              // integer division, UpdateSigma.h:49
  }
  e->n_points = (double)N; e->sum_half = sum_half;
  {
    std::vector<double> nih((size_t)e->ld, 0.0);
    for (int i = 0; i < n; i++) nih[i] = (double)(c->off[i + 1] - c->off[i]);
    CU(cudaMalloc(&e->ni, (size_t)e->ld * 8));
    CU(cudaMemcpy(e->ni, nih.data(), (size_t)e->ld * 8, cudaMemcpyHostToDevice));
  }
  std::vector<int64_t> off(c->off, c->off + n + 1);
  for (auto& o : off) o -= c->off[0];
  int64_t* d_off = nullptr; double *d_y = nullptr, *d_t = nullptr, *d_B = nullptr, *d_kn = nullptr; int* d_bw = nullptr;
  CU(cudaMalloc(&d_off, (size_t)(n + 1) * 8)); CU(cudaMalloc(&d_y, (size_t)N * 8));
  CU(cudaMemcpyAsync(d_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(d_y, c->y + c->off[0], (size_t)N * 8, cudaMemcpyHostToDevice, e->stream));
  bf::RaggedPrepArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  std::vector<double> kn;
  if (c->B) {
    CU(cudaMalloc(&d_B, (size_t)N * P * 8));
    CU(cudaMemcpyAsync(d_B, c->B + (size_t)c->off[0] * P, (size_t)N * P * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMalloc(&d_bw, 4)); CU(cudaMemsetAsync(d_bw, 0, 4, e->stream));
    if (bf::launch_band_width(d_B, N, P, d_bw, e->stream)) return fail("band width kernel launch failed");
    CU(cudaMemcpyAsync(&e->bw, d_bw, 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->bw < 1) e->bw = 1;
    pa.Brows = d_B;
  } else {
    if (!c->t || (c->n_internal > 0 && !c->internal_knots)) return fail("bfmmm_create: neither B nor a spline description (t, knots) was given");
    if (c->n_internal + c->degree + 1 != P) return fail("bfmmm_create: P != n_internal + degree + 1");
    int nk = c->n_internal + 2 * (c->degree + 1);
    kn.resize(nk);
    for (int i = 0; i <= c->degree; i++) { kn[i] = c->boundary[0]; kn[nk - 1 - i] = c->boundary[1]; }
    for (int i = 0; i < c->n_internal; i++) kn[c->degree + 1 + i] = c->internal_knots[i];
    CU(cudaMalloc(&d_t, (size_t)N * 8)); CU(cudaMalloc(&d_kn, (size_t)nk * 8));
    CU(cudaMemcpyAsync(d_t, c->t + c->off[0], (size_t)N * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_kn, kn.data(), (size_t)nk * 8, cudaMemcpyHostToDevice, e->stream));
    e->bw = c->degree + 1;
    pa.t = d_t; pa.knots = d_kn; pa.n_knots = nk; pa.degree = c->degree;
  }
  if (e->bw > bf::BWMAX) return fail("bfmmm_create: ragged grids support a basis band width (degree + 1) of at most 6");
  if (e->bw > P) e->bw = P;
  CU(cudaMalloc(&e->Hh, (size_t)e->ld * P * 8));
  CU(cudaMalloc(&e->Gl, (size_t)e->ld * P * e->bw * 8));
  CU(cudaMemsetAsync(e->Hh, 0, (size_t)e->ld * P * 8, e->stream));
  CU(cudaMemsetAsync(e->Gl, 0, (size_t)e->ld * P * e->bw * 8, e->stream));
  pa.n = n; pa.ld = e->ld; pa.P = P; pa.bw = e->bw; pa.i_begin = 0; pa.off = d_off; pa.y = d_y;
  pa.C = e->Ct; pa.H = e->Hh; pa.Gl = e->Gl; pa.rss = e->rss;
  int rc = bf::launch_ragged_prep(pa, e->stream);
  if (rc) return fail("ragged prep kernel launch failed (P <= 64 and band width <= 8 supported) rc=" + std::to_string(rc));
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(d_off); cudaFree(d_y); cudaFree(d_t); cudaFree(d_B); cudaFree(d_kn); cudaFree(d_bw);
  // statistics: pair cross-Gram
  e->npairs = e->q * (e->q + 1) / 2;
  CU(cudaMalloc(&e->rs_partials, bf::ragged_stats_partial_doubles(P, e->bw, e->q, e->sm_count) * 8));
  return 0;
}

}  // namespace

extern "C" {

const char* bfmmm_last_error(void) { return g_err.c_str(); }
int64_t bfmmm_launch_count(void) { return (int64_t)bf::g_launch_count.load(); }

int bfmmm_create(const bfmmm_config* c, bfmmm_engine** out) {
  if (!c || !out) return fail("bfmmm_create: null argument");
  *out = nullptr;
  if (c->n <= 0 || c->K < 2 || c->K > 6 || c->M < 1 || c->M > 6 || c->P < 1)
    return fail("bfmmm_create: unsupported shape (need n > 0, 2 <= K <= 6, 1 <= M <= 6, P >= 1)");
  if (c->D < 0 || c->D > bf::DMAX) return fail("bfmmm_create: D must be in [0, 4]");
  if (c->D > 0 && !c->X) return fail("bfmmm_create: D > 0 but X is NULL");
  if (!c->y) return fail("bfmmm_create: y is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("bfmmm_create: no CUDA device (this engine has no CPU fallback)");
  if (c->device < 0 || c->device >= ndev) return fail("bfmmm_create: bad device ordinal");
  CU(cudaSetDevice(c->device));
  bfmmm_engine* e = new bfmmm_engine();
  e->model = c->model; e->n = c->n; e->K = c->K; e->P = c->P; e->M = c->M; e->D = c->D; e->device = c->device;
  e->identity = (c->model == BFMMM_MULTIVARIATE);
  e->common = e->identity || c->common_grid;
  e->T = e->identity ? c->P : c->T;
  e->global_offset = c->global_offset;
  e->ld = (c->n + 63) & ~63;      // padded with zero functions: 8-function DMMA chunks, 64-function TMA stages
  e->q = c->K * (1 + c->D) * (1 + c->M);
  e->QS = (e->q + 1) & ~1;
  e->ragged = !e->common;
  if (e->common && !e->identity && e->T < 1) { delete e; return fail("bfmmm_create: common grid needs T >= 1"); }
  if (e->ragged && (e->q > 40 || e->P > 64)) { delete e; return fail("bfmmm_create: ragged grids support P <= 64 and q = K(1+D)(1+M) <= 40"); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c->device);
  e->sm_count = prop.multiProcessorCount;
  {
    // the pass kernels stage the whitened globals (P4 x QS doubles) in shared memory next to ~6 KB of static tables
    const size_t need = (size_t)((e->P + 3) & ~3) * e->QS * 8 + 6 * 1024;
    if (need > (size_t)prop.sharedMemPerBlockOptin) {
      const int q = e->q;
      delete e;
      return fail("bfmmm_create: P * q = " + std::to_string(c->P) + " * " + std::to_string(q) +
                  " doubles of global coefficients do not fit the " + std::to_string(prop.sharedMemPerBlockOptin / 1024) +
                  " KB of shared memory per block (need P4 * q * 8 + 6 KB; q = K (1 + D)(1 + M))");
    }
  }
  auto bail = [&](int) { free_all(e); return 1; };
#define CUE(x)                                                                  \
  do {                                                                          \
    cudaError_t _e = (x);                                                       \
    if (_e != cudaSuccess) { fail(std::string(#x) + ": " + cudaGetErrorString(_e)); return bail(1); } \
  } while (0)
  {
    // Stream priorities: the side stream (ahead-of-time Z proposals, z_propose_kernel) OUTRANKS the engine's stream.  The
    // proposal is launched into the gap in which the device waits for the host's block draws; when the SSR pass arrives
    // before it has finished, its remaining blocks go first and the pass then runs undisturbed.  The other way round the
    // persistent SSR pass takes the SMs as the proposal's blocks retire, the proposal finishes behind it and the next Z
    // step waits (measured on 2 GPUs: no gain at all from running ahead).  BFMMM_SIDE_PRIO=low restores that order.
    int least = 0, greatest = 0;
    CUE(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const char* pr = std::getenv("BFMMM_SIDE_PRIO");
    const bool side_low = pr && pr[0] == 'l';
    e->prio_side = side_low ? least : greatest;
    CUE(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, side_low ? greatest : least));
  }
  const size_t ld = e->ld;
  e->P4 = (e->P + 3) & ~3;
  CUE(cudaMalloc(&e->Ct, ld * e->P4 * 8));
  CUE(cudaMalloc(&e->rss, ld * 8));
  CUE(cudaMalloc(&e->Z, ld * e->K * 8));
  CUE(cudaMalloc(&e->chi, ld * e->M * 8));
  if (e->D) CUE(cudaMalloc(&e->X, ld * e->D * 8));
  CUE(cudaMalloc(&e->glob, (size_t)e->P4 * e->QS * 8));
  CUE(cudaMalloc(&e->draws, ld * (std::max(e->K + 1, e->M)) * 8));
  e->stats_len = e->K + 3 + (int64_t)e->q * e->q + (int64_t)e->P * e->q;
  if (e->ragged) e->stats_len += (int64_t)(e->q * (e->q + 1) / 2) * bf::BWMAX * e->P;   // upper bound (bw <= BWMAX)
  CUE(cudaMalloc(&e->stats, e->stats_len * 8));
  e->pass_blocks = std::min(bf::pass_grid(e->ld, 1), e->sm_count * 16);
  CUE(cudaMalloc(&e->partials, (size_t)e->pass_blocks * bf::RED_MAX * 8));
  e->st_blocks = bf::stats_blocks(e->sm_count);
  CUE(cudaMalloc(&e->st_partials, bf::stats_partial_doubles(e->P, e->q, e->st_blocks) * 8));
  CUE(cudaMalloc(&e->ticket, 4));
  CUE(cudaMemsetAsync(e->ticket, 0, 4, e->stream));
  CUE(cudaMalloc(&e->st_ticket, 4));
  CUE(cudaMemsetAsync(e->st_ticket, 0, 4, e->stream));
  CUE(cudaMemsetAsync(e->Ct, 0, ld * e->P4 * 8, e->stream));
  CUE(cudaMemsetAsync(e->rss, 0, ld * 8, e->stream));
  CUE(cudaMemsetAsync(e->Z, 0, ld * e->K * 8, e->stream));
  CUE(cudaMemsetAsync(e->chi, 0, ld * e->M * 8, e->stream));
  CUE(cudaMemsetAsync(e->stats, 0, e->stats_len * 8, e->stream));
  CUE(cudaMemsetAsync(e->draws, 0, ld * (std::max(e->K + 1, e->M)) * 8, e->stream));
  if (e->D) CUE(cudaMemsetAsync(e->X, 0, ld * e->D * 8, e->stream));
  for (int i = 0; i < N_STAGE; i++) {
    CUE(cudaMallocHost(&e->h_stage[i], (size_t)e->P4 * e->QS * 8));
    CUE(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
  }
  CUE(cudaHostAlloc(&e->h_stats, e->stats_len * 8, cudaHostAllocMapped));
  CUE(cudaHostGetDevicePointer(&e->h_stats_dev, e->h_stats, 0));
  CUE(cudaHostAlloc(&e->h_sig, 4 * 8, cudaHostAllocMapped));
  CUE(cudaHostGetDevicePointer(&e->h_sig_dev, e->h_sig, 0));
  e->h_sig[0] = e->h_sig[1] = e->h_sig[2] = e->h_sig[3] = 0;
  CUE(cudaMalloc(&e->sigma_dev, 8));
  if (e->identity) {
    e->G.assign((size_t)e->P * e->P, 0.0);
    e->L.assign((size_t)e->P * e->P, 0.0);
    for (int p = 0; p < e->P; p++) { e->G[(size_t)p * e->P + p] = 1; e->L[(size_t)p * e->P + p] = 1; }
    e->Pc = e->P;
    if (upload_cols(e, e->Ct, c->y, e->P)) return bail(1);      // c~_i = y_i, rss_i = 0
    e->n_points = (double)e->n * e->P;
    e->sum_half = (double)(((int64_t)e->n * e->P) / 2);          // y_obs.n_elem / 2, UpdateSigma.h:150
  } else if (e->ragged) {
    e->G.assign((size_t)e->P * e->P, 0.0);
    e->L.assign((size_t)e->P * e->P, 0.0);
    for (int p = 0; p < e->P; p++) e->L[(size_t)p * e->P + p] = 1;      // no common whitening
    e->Pc = e->P;
    if (project_ragged(e, c)) return bail(1);
    e->stats_len = e->K + 3 + (int64_t)e->q * e->q + (int64_t)e->P * e->q + (int64_t)e->npairs * e->bw * e->P;
  } else {
    if (build_basis(e, c)) return bail(1);
    e->hbL = e->P - 1;
    if (e->Pc == e->P) {                       // Cholesky factor of a banded Gram matrix: same band
      int hb = 0;
      for (int p = 0; p < e->P; p++)
        for (int r = p; r < e->P; r++)
          if (e->L[(size_t)p * e->P + r] != 0.0 && r - p > hb) hb = r - p;
      e->hbL = hb;
    }
    if (project_common(e, c)) return bail(1);
    e->n_points = (double)e->n * (double)e->T;
    e->sum_half = (double)e->n * (double)(e->T / 2);             // sum_i floor(n_i / 2), UpdateSigma.h:49
  }
  if (e->D && upload_cols(e, e->X, c->X, e->D)) return bail(1);
  e->z_split = !e->ragged && e->K >= 2 && e->K <= 6 && !std::getenv("BFMMM_Z_FUSED");
  e->mom_enabled = !e->ragged && e->D == 0 && e->M >= 1 && e->M <= 6 && e->K >= 2 && e->K <= 6 && !std::getenv("BFMMM_NO_MOMENTS");
  e->tma.valid = 0;
  if (!e->ragged) bf::stats_tma_setup(&e->tma, e->Ct, e->Z, e->chi, e->X, e->ld, e->Pc, e->K, e->M, e->D, e->q);
  CUE(cudaStreamSynchronize(e->stream));
#undef CUE
  *out = e;
  return 0;
}

void bfmmm_destroy(bfmmm_engine* e) {
  if (e && std::getenv("BFMMM_DEBUG") && (e->n_prop_ahead || e->n_prop_own))
    std::fprintf(stderr, "[bfmmm debug] Z steps: %lld with an ahead-of-time proposal, %lld with their own\n", e->n_prop_ahead, e->n_prop_own);
  free_all(e);
}

int bfmmm_get_basis(bfmmm_engine* e, double* B_out) {
  if (!e || e->identity) return fail("bfmmm_get_basis: no basis for this model");
  std::copy(e->B.begin(), e->B.end(), B_out);
  return 0;
}
int bfmmm_engine_dims(bfmmm_engine* e, int32_t* dims) {
  if (!e || !dims) return fail("null argument");
  dims[0] = e->n; dims[1] = e->K; dims[2] = e->P; dims[3] = e->M; dims[4] = e->D; dims[5] = e->model;
  dims[6] = e->ragged ? 1 : 0; dims[7] = e->bw;
  return 0;
}
int bfmmm_counts(bfmmm_engine* e, double* sum_half, double* n_points) {
  if (!e) return fail("null engine");
  if (sum_half) *sum_half = e->sum_half;
  if (n_points) *n_points = e->n_points;
  return 0;
}
int bfmmm_get_gram(bfmmm_engine* e, double* G) {
  if (!e) return fail("null engine");
  if (e->ragged) return fail("bfmmm_get_gram: ragged grids have one Gram matrix per function (use bfmmm_suffstats_ragged)");
  std::copy(e->G.begin(), e->G.end(), G);
  return 0;
}

static void z_written(bfmmm_engine* e);
// BFMMM_DEBUG: reports (and clears) a CUDA error some earlier unchecked call left behind
static void dbg_stale(const char* where) {
  static const bool on = std::getenv("BFMMM_DEBUG") != nullptr;
  if (!on) return;
  cudaError_t pe = cudaGetLastError();
  if (pe != cudaSuccess) std::fprintf(stderr, "[bfmmm debug] stale CUDA error at %s: %s\n", where, cudaGetErrorString(pe));
}
int bfmmm_set_state(bfmmm_engine* e, const double* Z, const double* chi) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (Z && upload_cols(e, e->Z, Z, e->K)) return 1;
  if (Z) z_written(e);
  if (chi && upload_cols(e, e->chi, chi, e->M)) return 1;
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
int bfmmm_get_state(bfmmm_engine* e, double* Z, double* chi) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (Z && download_cols(e, Z, e->Z, e->K)) return 1;
  if (chi && download_cols(e, chi, e->chi, e->M)) return 1;
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

// Overlapped read-back: device-side snapshot on the compute stream, host transfer on a copy stream.
int bfmmm_get_state_begin(bfmmm_engine* e, double* Z, double* chi) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (!e->copy_stream) {
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&e->ev_snap, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->ev_copied, cudaEventDisableTiming));
    CU(cudaMalloc(&e->rbZ, (size_t)e->ld * e->K * 8));
    CU(cudaMalloc(&e->rbChi, (size_t)e->ld * e->M * 8));
  }
  if (e->rb_pending) CU(cudaStreamWaitEvent(e->stream, e->ev_copied, 0));   // the staging buffers are being read
  if (Z) CU(cudaMemcpyAsync(e->rbZ, e->Z, (size_t)e->ld * e->K * 8, cudaMemcpyDeviceToDevice, e->stream));
  if (chi) CU(cudaMemcpyAsync(e->rbChi, e->chi, (size_t)e->ld * e->M * 8, cudaMemcpyDeviceToDevice, e->stream));
  CU(cudaEventRecord(e->ev_snap, e->stream));
  CU(cudaStreamWaitEvent(e->copy_stream, e->ev_snap, 0));
  if (Z) CU(cudaMemcpy2DAsync(Z, (size_t)e->n * 8, e->rbZ, (size_t)e->ld * 8, (size_t)e->n * 8, e->K, cudaMemcpyDeviceToHost, e->copy_stream));
  if (chi) CU(cudaMemcpy2DAsync(chi, (size_t)e->n * 8, e->rbChi, (size_t)e->ld * 8, (size_t)e->n * 8, e->M, cudaMemcpyDeviceToHost, e->copy_stream));
  CU(cudaEventRecord(e->ev_copied, e->copy_stream));
  e->rb_pending = true;
  return 0;
}
int bfmmm_get_state_wait(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  if (!e->rb_pending) return 0;
  CU(cudaSetDevice(e->device));
  CU(cudaEventSynchronize(e->ev_copied));
  e->rb_pending = false;
  return 0;
}

// rows [i0, i0 + count) of Z (count x K) and chi (count x M), column-major with leading dimension count:
// a cheap read-back of a few functions (chain monitoring / ESS) instead of the whole state
int bfmmm_get_state_rows(bfmmm_engine* e, int64_t i0, int64_t count, double* Z, double* chi) {
  if (!e) return fail("null engine");
  if (i0 < 0 || count < 0 || i0 + count > e->n) return fail("bfmmm_get_state_rows: range outside the shard");
  CU(cudaSetDevice(e->device));
  if (Z) CU(cudaMemcpy2DAsync(Z, (size_t)count * 8, e->Z + i0, (size_t)e->ld * 8, (size_t)count * 8, e->K, cudaMemcpyDeviceToHost, e->stream));
  if (chi) CU(cudaMemcpy2DAsync(chi, (size_t)count * 8, e->chi + i0, (size_t)e->ld * 8, (size_t)count * 8, e->M, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

// whitened coefficient of feature f: c~ = L' c
int bfmmm_set_globals(bfmmm_engine* e, const double* nu, const double* Phi, const double* eta,
                      const double* xi, double sigma_sq) {
  if (!e) return fail("null engine");
  if (!nu || (e->M > 0 && !Phi) || (e->D > 0 && (!eta || !xi))) return fail("bfmmm_set_globals: null block");
  if (!(sigma_sq > 0)) return fail("bfmmm_set_globals: sigma_sq must be positive");
  CU(cudaSetDevice(e->device));
  const int K = e->K, P = e->P, M = e->M, D = e->D, QS = e->QS;
  e->mom_valid = false;
  int slot = e->stage_next;
  e->stage_next = (slot + 1) % N_STAGE;
  CU(cudaEventSynchronize(e->ev_stage[slot]));
  double* h = e->h_stage[slot];
  std::vector<double> cvec(P);
  for (int f = 0; f < e->q; f++) {
    int dd = f % (1 + D), km = f / (1 + D), mm = km % (M + 1), k = km / (M + 1);
    for (int p = 0; p < P; p++) {
      double v;
      if (mm == 0 && dd == 0) v = nu[(size_t)p * K + k];
      else if (mm == 0) v = eta[((size_t)k * D + (dd - 1)) * P + p];
      else if (dd == 0) v = Phi[((size_t)(mm - 1) * P + p) * K + k];
      else v = xi[(size_t)k * P * D * M + ((size_t)(mm - 1) * D + (dd - 1)) * P + p];
      cvec[p] = v;
    }
    if (e->identity || e->ragged) {
      for (int p = 0; p < P; p++) h[(size_t)p * QS + f] = cvec[p];
    } else {
      for (int p = 0; p < e->Pc; p++) {           // (L' c)[p] = sum_r L[r][p] c[r]
        double s = 0;
        const int r_hi = (e->Pc == P) ? std::min(P - 1, p + e->hbL) : P - 1;
        for (int r = (e->Pc == P ? p : 0); r <= r_hi; r++) s += e->L[(size_t)p * P + r] * cvec[r];
        h[(size_t)p * QS + f] = s;
      }
    }
  }
  for (int p = 0; p < e->Pc; p++)
    for (int f = e->q; f < QS; f++) h[(size_t)p * QS + f] = 0.0;
  for (size_t idx = (size_t)e->Pc * QS; idx < (size_t)e->P4 * QS; idx++) h[idx] = 0.0;     // padding rows
  CU(cudaMemcpyAsync(e->glob, h, (size_t)e->P4 * QS * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaEventRecord(e->ev_stage[slot], e->stream));
  e->sigma_sq = sigma_sq;
  return 0;
}

static void fill_pass(bfmmm_engine* e, bf::PassArgs& a, double beta) {
  std::memset(&a, 0, sizeof(a));
  e->mirror_valid = false; e->stats_exchanged = false;      // every pass kernel writes a slot of the statistics header
  a.n = e->n; a.ld = e->ld; a.P = e->Pc; a.D = e->D; a.QS = e->QS; a.P4 = (e->Pc + 3) & ~3;
  a.sm_count = e->sm_count; a.max_blocks = e->pass_blocks;
  a.Ct = e->Ct; a.Gl = e->Gl; a.bw = e->bw; a.rss = e->rss; a.Z = e->Z; a.chi = e->chi; a.X = e->X; a.glob = e->glob;
  a.sigma_sq = e->sigma_sq; a.beta = beta;
  a.key = e->key; a.iteration = e->iteration; a.global_offset = (uint64_t)e->global_offset;
  bf::philox_round_keys(e->key, a.rk);
  a.partials = e->partials; a.ticket = e->ticket;
}

// digamma and trigamma at x > 0: upward recurrence to x >= 12, then the asymptotic series
static void polygamma01(double x, double& digam, double& trigam) {
  double d = 0, t = 0;
  while (x < 12.0) { d -= 1.0 / x; t += 1.0 / (x * x); x += 1.0; }
  const double r = 1.0 / x, r2 = r * r;
  d += std::log(x) - 0.5 * r - r2 * (1.0 / 12 - r2 * (1.0 / 120 - r2 * (1.0 / 252 - r2 * (1.0 / 240 - r2 * (1.0 / 132)))));
  t += r + 0.5 * r2 + r2 * r * (1.0 / 6 - r2 * (1.0 / 30 - r2 * (1.0 / 42 - r2 * (1.0 / 30 - r2 * (5.0 / 66)))));
  digam = d; trigam = t;
}

// Z was (or is being, on the engine's stream) written by something other than the Z step
static void z_written(bfmmm_engine* e) {
  e->mom_valid = false; e->prop_valid = false;
  if (e->ev_zdone) cudaEventRecord(e->ev_zdone, e->stream);
}
static void z_fill(bfmmm_engine* e, bf::PassArgs& a, const double* pi, double alpha3, double a_Z_PM, const double* zpar_dev) {
  a.zpar_dev = zpar_dev;
  a.alpha3 = alpha3; a.a_Z_PM = a_Z_PM; a.log_a_Z_PM = std::log(a_Z_PM); a.inv_a_Z_PM = 1.0 / a_Z_PM;
  double digam_a = 0;
  polygamma01(a_Z_PM, digam_a, a.trigam_a);
  a.c_tot = 1.0 - a.log_a_Z_PM + digam_a;
  for (int k = 0; k < e->K; k++) a.pi[k] = pi ? pi[k] : 0.0;
}
static int z_split_init(bfmmm_engine* e) {
  if (e->zprop) return 0;
  CU(cudaMalloc(&e->zprop, (size_t)e->ld * (e->K + 2) * 8));
  CU(cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, e->prio_side));
  CU(cudaHostAlloc(&e->h_slz, 16 * 8, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer(&e->h_slz_dev, e->h_slz, 0));
  CU(cudaEventCreateWithFlags(&e->ev_slz, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->ev_zdone, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->ev_queued, cudaEventDisableTiming));
  CU(cudaEventCreate(&e->ev_k1a));
  CU(cudaEventCreate(&e->ev_k1b));
  CU(cudaEventCreateWithFlags(&e->ev_prop, cudaEventDisableTiming));
  return 0;
}

static int z_launch(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta, bool injected,
                    bool dump_draws, const double* zpar_dev = nullptr) {
  dbg_stale("z_launch entry");
  bf::PassArgs a;
  fill_pass(e, a, beta);
  e->mom_valid = false;
  z_fill(e, a, pi, alpha3, a_Z_PM, zpar_dev);
#ifdef BF_TUNE_V
  static const bool force_inject = std::getenv("BFMMM_Z_INJECT") != nullptr;   // tuning: time the step without the RNG
  injected = injected || force_inject;
#endif
  if (injected) { a.gam = e->draws; a.u = e->draws + (size_t)e->K * e->ld; }
  if (dump_draws) a.draws_out = e->draws;
  a.acc_out = e->acc_dbg;
  a.out = e->stats + e->off_slz(); a.n_out = e->K + 1;
  int rc;
  if (e->z_split) {
    if (z_split_init(e)) return 1;
    bool ahead = e->prop_valid && !injected && !dump_draws && !zpar_dev && pi && e->prop_alpha3 == alpha3 && e->prop_a == a_Z_PM &&
                 e->prop_key == e->key && e->prop_iter == e->iteration;
    for (int k = 0; ahead && k < e->K; k++) ahead = (e->prop_pi[k] == pi[k]);
    e->prop_valid = false;
    if (ahead) {
      e->n_prop_ahead++;
      CU(cudaStreamWaitEvent(e->stream, e->ev_prop, 0));          // made ahead of time on the side stream
    } else {
      CU(cudaStreamWaitEvent(e->stream, e->ev_prop, 0));          // a stale proposal may still be writing the buffer
      e->n_prop_own++;
      a.zprop_out = e->zprop;
      CU(cudaEventRecord(e->ev_k1a, e->stream));
      rc = bf::launch_z_propose(a, e->K, e->stream);
      if (rc) return fail("z proposal kernel launch failed rc=" + std::to_string(rc));
      CU(cudaEventRecord(e->ev_k1b, e->stream));
      e->k1_timed = true;
    }
    a.zprop = e->zprop; a.zprop_out = nullptr;
    a.gam = nullptr; a.u = nullptr; a.draws_out = nullptr;
    rc = bf::launch_z_accept(a, e->K, e->M, e->stream);
    CU(cudaEventRecord(e->ev_zdone, e->stream));
  } else {
    rc = e->ragged ? bf::launch_z_ragged(a, e->K, e->M, e->stream) : bf::launch_z(a, e->K, e->M, e->stream);
  }
  if (rc) return fail("z kernel launch failed rc=" + std::to_string(rc));
  dbg_stale("z_launch exit");
  return 0;
}

int bfmmm_update_z_async(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta) {
  if (!e || !pi) return fail("null argument");
  CU(cudaSetDevice(e->device));
  return z_launch(e, pi, alpha3, a_Z_PM, beta, false, false);
}

int bfmmm_update_z(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta,
                   const double* gam, const double* u, double* sum_log_Z, int64_t* n_accept) {
  if (!e || !pi) return fail("null argument");
  if ((gam == nullptr) != (u == nullptr)) return fail("bfmmm_update_z: gam and u must both be given or both be NULL");
  CU(cudaSetDevice(e->device));
  if (gam) {
    if (upload_cols(e, e->draws, gam, e->K)) return 1;
    if (upload_cols(e, e->draws + (size_t)e->K * e->ld, u, 1)) return 1;
  }
  if (z_launch(e, pi, alpha3, a_Z_PM, beta, gam != nullptr, false)) return 1;
  CU(cudaMemcpyAsync(e->h_stats, e->stats + e->off_slz(), (e->K + 1) * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (sum_log_Z) std::copy(e->h_stats, e->h_stats + e->K, sum_log_Z);
  if (n_accept) *n_accept = (int64_t)std::llround(e->h_stats[e->K]);
  return 0;
}

static int chi_launch(bfmmm_engine* e, double beta, bool injected, const double* sigma_dev = nullptr, bool dump_draws = false) {
  bf::PassArgs a;
  fill_pass(e, a, beta);
  if (e->sigma_armed) { a.sigma_dev = e->sigma_dev; e->sigma_armed = false; }
  if (sigma_dev) {
    a.sigma_dev = sigma_dev;
    // device-resident sweep: the shrinkage-prior kernel of the side stream may still hold 4096 registers of one SM,
    // and this kernel's blocks take all 65536: one block fewer keeps the whole grid resident from the start
    a.grid_reserve = 1;
  }
  if (injected) a.eps = e->draws;
  if (dump_draws) a.draws_out = e->draws;
  a.out = e->stats + e->off_ssr_after(); a.n_out = 1;
  // the SSR pass that preceded this step left the moments of the same (Z, globals): draw from them, no second data pass
  int rc = e->mom_valid ? bf::launch_chi_draw(a, e->K, e->M, e->mom, e->stream)
           : e->ragged  ? bf::launch_chi_ragged(a, e->K, e->M, e->stream)
                        : bf::launch_chi(a, e->K, e->M, e->stream);
  if (rc) return fail("chi kernel launch failed rc=" + std::to_string(rc));
  return 0;
}
int bfmmm_update_chi_async(bfmmm_engine* e, double beta) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  return chi_launch(e, beta, false);
}
int bfmmm_update_chi(bfmmm_engine* e, double beta, const double* eps, double* ssr_after) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (eps && upload_cols(e, e->draws, eps, e->M)) return 1;
  if (chi_launch(e, beta, eps != nullptr)) return 1;
  CU(cudaMemcpyAsync(e->h_stats, e->stats + e->off_ssr_after(), 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (ssr_after) *ssr_after = e->h_stats[0];
  return 0;
}

static int ssr_launch(bfmmm_engine* e, const bf::SigmaTail* tail) {
  bf::PassArgs a;
  fill_pass(e, a, 1.0);
  a.out = e->stats + e->off_ssr(); a.n_out = 1;
  int rc;
  if (e->mom_enabled) {          // common basis, no covariates: the pass also leaves the chi step's moments
    if (!e->mom) CU(cudaMalloc(&e->mom, (size_t)e->ld * (e->M + 1) * 8));
    rc = bf::launch_moments(a, e->K, e->M, e->mom, tail, e->stream);
    e->mom_valid = (rc == 0);
  } else {
    rc = e->ragged ? bf::launch_ssr_ragged(a, e->K, e->M, e->stream) : bf::launch_ssr(a, e->K, e->M, e->stream);
  }
  if (rc) return fail("ssr kernel launch failed rc=" + std::to_string(rc));
  return 0;
}
int bfmmm_ssr_async(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  return ssr_launch(e, nullptr);
}
int bfmmm_ssr(bfmmm_engine* e, double* ssr, double* sum_half, double* n_points) {
  if (bfmmm_ssr_async(e)) return 1;
  CU(cudaMemcpyAsync(e->h_stats, e->stats + e->off_ssr(), 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (ssr) *ssr = e->h_stats[0];
  if (sum_half) *sum_half = e->sum_half;
  if (n_points) *n_points = e->n_points;
  return 0;
}

// ---- post-processing: per-function marginal log-likelihood and CPO (calcLikelihoodCPO, CalculateLikelihood.h:344-385)
static int mloglik_launch(bfmmm_engine* e, bool accumulate) {
  if (!e->logl) CU(cudaMalloc(&e->logl, (size_t)e->ld * 8));
  if (accumulate && !e->cpo_m) {
    CU(cudaMalloc(&e->cpo_m, (size_t)e->ld * 8));
    CU(cudaMalloc(&e->cpo_s, (size_t)e->ld * 8));
  }
  bf::PassArgs a;
  fill_pass(e, a, 1.0);
  a.ni = e->ni; a.npts_common = e->identity ? (double)e->P : (double)e->T;
  a.logl_out = e->logl;
  if (accumulate) { a.cpo_m = e->cpo_m; a.cpo_s = e->cpo_s; a.cpo_first = e->cpo_count == 0 ? 1 : 0; }
  int rc = bf::launch_mloglik(a, e->K, e->M, e->ragged, e->stream);
  if (rc) return fail("marginal log-likelihood kernel launch failed rc=" + std::to_string(rc));
  if (accumulate) e->cpo_count++;
  return 0;
}
// log p(y_i | Z_i, globals, sigma^2) with chi_i integrated out, for the current state and the globals last pushed
int bfmmm_marginal_loglik(bfmmm_engine* e, double* logl /* n */) {
  if (!e || !logl) return fail("null argument");
  CU(cudaSetDevice(e->device));
  if (mloglik_launch(e, false)) return 1;
  CU(cudaMemcpyAsync(logl, e->logl, (size_t)e->n * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
int bfmmm_cpo_reset(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  e->cpo_count = 0;
  return 0;
}
// adds the current state (one retained iteration) to the running harmonic mean; asynchronous
int bfmmm_cpo_accumulate(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  return mloglik_launch(e, true);
}
// CPO_i = log L - log sum_l exp(-logl_il) over the L accumulated iterations (log scale unless log_scale == 0)
int bfmmm_cpo_get(bfmmm_engine* e, double* cpo /* n */, int log_scale) {
  if (!e || !cpo) return fail("null argument");
  if (e->cpo_count == 0) return fail("bfmmm_cpo_get: nothing accumulated");
  CU(cudaSetDevice(e->device));
  std::vector<double> m((size_t)e->n), sv((size_t)e->n);
  CU(cudaMemcpyAsync(m.data(), e->cpo_m, (size_t)e->n * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(sv.data(), e->cpo_s, (size_t)e->n * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  const double lL = std::log((double)e->cpo_count);
  for (int i = 0; i < e->n; i++) {
    const double v = lL - (m[i] + std::log(sv[i]));
    cpo[i] = log_scale ? v : std::exp(v);
  }
  return 0;
}

// sigma^2 drawn on the device behind the SSR pass (see sigma_draw_kernel); arms the next chi launch
int bfmmm_sigma_draw_async(bfmmm_engine* e, double a, double scale_ssr, double beta0, uint64_t key, uint64_t iteration,
                           uint32_t purpose) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  e->sig_seq += 1.0;
  if (bf::launch_sigma_draw(e->stats + e->off_ssr(), a, scale_ssr, beta0, key, iteration, purpose, e->sigma_dev, e->h_sig_dev,
                            e->sig_seq, e->stream))
    return fail("sigma kernel launch failed");
  e->sigma_armed = true;
  return 0;
}
// waits (polling the mapped sequence number, not the stream: later kernels may already be queued) for the
// device draw and returns SSR and sigma^2; sigma^2 becomes the engine's value for the launches that follow
int bfmmm_sigma_wait(bfmmm_engine* e, double* ssr, double* sigma_sq) {
  if (!e) return fail("null engine");
  volatile double* h = e->h_sig;
  for (long long spin = 0; h[2] != e->sig_seq; spin++) {
    if ((spin & 0xfffff) == 0xfffff && cudaStreamQuery(e->stream) != cudaErrorNotReady && h[2] != e->sig_seq) {
      cudaError_t err = cudaGetLastError();
      return fail(std::string("bfmmm_sigma_wait: the device draw did not arrive: ") + cudaGetErrorString(err));
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (ssr) *ssr = h[0];
  if (sigma_sq) *sigma_sq = h[1];
  e->sigma_sq = h[1];
  return 0;
}

int bfmmm_suffstats_async(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  dbg_stale("suffstats entry");
  bf::StatsArgs a;
  a.n = e->n; a.ld = e->ld; a.P = e->Pc; a.K = e->K; a.M = e->M; a.D = e->D; a.q = e->q;
  a.tma = (!e->ragged && e->tma.valid) ? &e->tma : nullptr;
  a.Ct = e->ragged ? e->Hh : e->Ct; a.Z = e->Z; a.chi = e->chi; a.X = e->X; a.partials = e->st_partials;
  a.WtW = e->stats + e->off_wtw(); a.CtW = e->stats + e->off_ctw(); a.blocks = e->st_blocks;
  // Epilogue of the final reduction.  On several GPUs with the peer-memory exchange installed, the last block of the
  // final reduction exchanges the buffer and stores the totals into the mapped host copy (one launch instead of three).
  // On one GPU the same trick (BFMMM_STATS_EPILOGUE=1) was measured SLOWER than the separate copy kernel (0.277 against
  // 0.273 ms per sweep: fence + ticket + a serial tail block cost more than a kernel boundary), so it is off.
  // Ragged grids append the pair cross-Gram band behind this pass: they keep the separate exchange and read-back.
  static const bool no_epilogue = std::getenv("BFMMM_NO_STATS_EPILOGUE") != nullptr;
  static const bool mirror_always = std::getenv("BFMMM_STATS_EPILOGUE") != nullptr;
  std::memset(&a.ep, 0, sizeof(a.ep));
  a.ep.stats = e->stats; a.ep.hdr = e->K + 3; a.ep.ticket = e->st_ticket; a.ep.world = 1;
  const bool epilogue = !e->ragged && !no_epilogue && (e->xchg_on || mirror_always);
  if (epilogue) a.ep.mirror = e->h_stats_dev;
  const bool exchange = epilogue && e->xchg_on;
  if (exchange) {
    a.ep.peers = e->xchg_peers; a.ep.rank = e->xchg_rank; a.ep.world = e->xchg_world; a.ep.cap = e->xchg_cap;
    a.ep.seq = ++*e->xchg_seq;
  }
  e->mirror_valid = false; e->stats_exchanged = false;
  int rc = bf::launch_stats(a, e->stream);
  if (rc) return fail("stats kernel launch failed rc=" + std::to_string(rc) + " (" + cudaGetErrorString((cudaError_t)rc) + ")");
  e->mirror_valid = epilogue; e->stats_exchanged = exchange;
  if (e->ragged) {
    bf::RaggedStatsArgs r;
    r.n = e->n; r.ld = e->ld; r.P = e->P; r.bw = e->bw; r.K = e->K; r.M = e->M; r.D = e->D; r.q = e->q; r.npairs = e->npairs;
    r.Gl = e->Gl; r.Z = e->Z; r.chi = e->chi; r.X = e->X; r.partials = e->rs_partials; r.Hb = e->stats + e->off_hb();
    rc = bf::launch_ragged_stats(r, e->sm_count, e->stream);
    if (rc) return fail("ragged stats kernel launch failed rc=" + std::to_string(rc));
  }
  return 0;
}

// B'Y'W = L * (C~'W): L is P x Pc (lower triangular when Pc == P), C~'W is Pc x q with leading dimension Pc
static void unwhiten(const bfmmm_engine* e, const double* CtW, double* BtYW) {
  const int P = e->P, Pc = e->Pc;
  for (int f = 0; f < e->q; f++)
    for (int r = 0; r < P; r++) {
      if (e->identity || e->ragged) { BtYW[(size_t)f * P + r] = CtW[(size_t)f * P + r]; continue; }
      double s = 0;
      const int kmax = (Pc == P) ? r + 1 : Pc;
      const int kmin = (Pc == P) ? std::max(0, r - e->hbL) : 0;
      for (int k = kmin; k < kmax; k++) s += e->L[(size_t)k * P + r] * CtW[(size_t)f * Pc + k];
      BtYW[(size_t)f * P + r] = s;
    }
}

int bfmmm_suffstats(bfmmm_engine* e, double* WtW, double* BtYW) {
  if (bfmmm_suffstats_async(e)) return 1;
  const int64_t len = (int64_t)e->q * e->q + (int64_t)e->P * e->q;
  CU(cudaMemcpyAsync(e->h_stats + e->off_wtw(), e->stats + e->off_wtw(), len * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (WtW) std::copy(e->h_stats + e->off_wtw(), e->h_stats + e->off_ctw(), WtW);
  if (BtYW) unwhiten(e, e->h_stats + e->off_ctw(), BtYW);
  return 0;
}

// device-side copy of the per-function state, so a rejected tempered transition
// (BFMMM.h:1631-1651 keeps the pre-transition slice) can be undone without a host round trip
int bfmmm_state_snapshot(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (!e->snapZ) CU(cudaMalloc(&e->snapZ, (size_t)e->ld * e->K * 8));
  if (!e->snapChi) CU(cudaMalloc(&e->snapChi, (size_t)e->ld * e->M * 8));
  CU(cudaMemcpyAsync(e->snapZ, e->Z, (size_t)e->ld * e->K * 8, cudaMemcpyDeviceToDevice, e->stream));
  CU(cudaMemcpyAsync(e->snapChi, e->chi, (size_t)e->ld * e->M * 8, cudaMemcpyDeviceToDevice, e->stream));
  return 0;
}
int bfmmm_state_restore(bfmmm_engine* e) {
  if (!e || !e->snapZ) return fail("bfmmm_state_restore: no snapshot");
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpyAsync(e->Z, e->snapZ, (size_t)e->ld * e->K * 8, cudaMemcpyDeviceToDevice, e->stream));
  z_written(e);
  CU(cudaMemcpyAsync(e->chi, e->snapChi, (size_t)e->ld * e->M * 8, cudaMemcpyDeviceToDevice, e->stream));
  return 0;
}

int bfmmm_suffstats_ragged(bfmmm_engine* e, double* WtW, double* BtYW, double* Hband) {
  if (!e || !e->ragged) return fail("bfmmm_suffstats_ragged: engine was not created with ragged grids");
  if (bfmmm_suffstats_async(e)) return 1;
  const int64_t len = e->stats_len - e->off_wtw();
  CU(cudaMemcpyAsync(e->h_stats + e->off_wtw(), e->stats + e->off_wtw(), len * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (WtW) std::copy(e->h_stats + e->off_wtw(), e->h_stats + e->off_ctw(), WtW);
  if (BtYW) std::copy(e->h_stats + e->off_ctw(), e->h_stats + e->off_hb(), BtYW);
  if (Hband) std::copy(e->h_stats + e->off_hb(), e->h_stats + e->stats_len, Hband);
  return 0;
}

// zeroes the post-chi SSR slot of the statistics buffer (after it has been summed over the shards and
// read, so that a later whole-buffer exchange does not add the global value once more)
int bfmmm_clear_ssr_after(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  e->mirror_valid = false;
  CU(cudaMemsetAsync(e->stats + e->off_ssr_after(), 0, 8, e->stream));
  return 0;
}

int bfmmm_seed(bfmmm_engine* e, uint64_t key, uint64_t iteration) {
  if (!e) return fail("null engine");
  e->key = key; e->iteration = iteration;
  return 0;
}

int bfmmm_stats_buffer_dev(bfmmm_engine* e, double** ptr, int64_t* len) {
  if (!e) return fail("null engine");
  *ptr = e->stats; *len = e->stats_len;
  e->mirror_valid = false;         // the caller may change the buffer (all-reduce hooks do)
  return 0;
}
int bfmmm_read_stats(bfmmm_engine* e, double* out, int64_t len) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (len > e->stats_len) len = e->stats_len;
  // SM-driven store into mapped host memory: does not queue behind an overlapped state transfer (the statistics pass's
  // epilogue has already stored the buffer there when nothing has touched it since)
  if (!e->mirror_valid && bf::launch_copy_to_host(e->stats, e->h_stats_dev, len, e->stream)) return fail("copy_to_host kernel launch failed");
  e->mirror_valid = false;         // the C~'W block is un-whitened in place below
  CU(cudaStreamSynchronize(e->stream));
  // the C~'W block is returned un-whitened (B'Y'W), like bfmmm_suffstats
  std::vector<double> tmp;
  if (len >= e->stats_len) {
    tmp.resize((size_t)e->P * e->q);
    unwhiten(e, e->h_stats + e->off_ctw(), tmp.data());
    std::copy(tmp.begin(), tmp.end(), e->h_stats + e->off_ctw());
  }
  std::copy(e->h_stats, e->h_stats + len, out);
  return 0;
}
int bfmmm_sync(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
void* bfmmm_stream(bfmmm_engine* e) { return e ? (void*)e->stream : nullptr; }

// ---- diagnostics used by the parity tests (declared in bfmmm_debug.h) ----
// per-function acceptance log-ratio of the last/next Z steps is written to an internal buffer
int bfmmm_debug_enable_acc(bfmmm_engine* e, int on) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (on && !e->acc_dbg) CU(cudaMalloc(&e->acc_dbg, (size_t)e->ld * 8));
  if (!on && e->acc_dbg) { cudaFree(e->acc_dbg); e->acc_dbg = nullptr; }
  return 0;
}
int bfmmm_debug_get_acc(bfmmm_engine* e, double* acc) {
  if (!e || !e->acc_dbg) return fail("acc diagnostics not enabled");
  CU(cudaMemcpyAsync(acc, e->acc_dbg, (size_t)e->n * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
// device-RNG Z step that also returns the draws it used (gam n x K, u n), so the oracle can be
// run on exactly the same draws
int bfmmm_debug_update_z_rng(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta,
                             double* gam_out, double* u_out) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (z_launch(e, pi, alpha3, a_Z_PM, beta, false, true)) return 1;
  if (download_cols(e, gam_out, e->draws, e->K)) return 1;
  if (download_cols(e, u_out, e->draws + (size_t)e->K * e->ld, 1)) return 1;
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
int bfmmm_debug_update_chi_rng(bfmmm_engine* e, double beta, double* eps_out) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (chi_launch(e, beta, false, nullptr, true)) return 1;
  if (download_cols(e, eps_out, e->draws, e->M)) return 1;
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
// the proposal half of the Z step alone, on the engine's stream (kernel timing; the result is not used)
int bfmmm_debug_z_propose(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM) {
  if (!e || !pi) return fail("null argument");
  if (!e->z_split) return fail("bfmmm_debug_z_propose: this engine runs the Z step as one kernel");
  CU(cudaSetDevice(e->device));
  if (z_split_init(e)) return 1;
  bf::PassArgs a;
  fill_pass(e, a, 1.0);
  z_fill(e, a, pi, alpha3, a_Z_PM, nullptr);
  e->prop_valid = false;
  CU(cudaStreamWaitEvent(e->stream, e->ev_prop, 0));
  a.zprop_out = e->zprop;
  CU(cudaEventRecord(e->ev_k1a, e->stream));          // the same event records as in the Z step, so that their cost
  if (bf::launch_z_propose(a, e->K, e->stream)) return fail("z proposal kernel launch failed");
  CU(cudaEventRecord(e->ev_k1b, e->stream));          // cancels in (whole step) - (this call)
  return 0;
}
// 1 when the next bfmmm_update_chi will draw from the moments the last SSR pass left (moments_kernels.cu)
int bfmmm_debug_moments_valid(bfmmm_engine* e) { return e && e->mom_valid ? 1 : 0; }
// the projected cache itself: C~ (n x P column-major) and rss (n)
int bfmmm_debug_get_cache(bfmmm_engine* e, double* Ct, double* rss) {
  if (!e) return fail("null engine");
  CU(cudaSetDevice(e->device));
  if (Ct && download_cols(e, Ct, e->Ct, e->Pc)) return 1;
  if (rss && download_cols(e, rss, e->rss, 1)) return 1;
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

}  // extern "C"

// ---- internal entry points of the device-resident sweep (engine_internal.h; not part of the C ABI)
// updateSigma in one launch: the SSR pass with the sigma^2 draw (and, over several GPUs, the one-slot exchange of the SSR
// through the peer-memory mailboxes) in its tail.  *done = 0 when the engine cannot do that (no moments pass for this
// model, or an exchange is needed and the peer-memory exchange is not installed): nothing was launched then and the caller
// runs bfmmm_ssr_async, its all-reduce hook and bfmmm_sigma_draw_async as before.
int bfmmm_ssr_sigma_async(bfmmm_engine* e, int need_exchange, double a_shape, double scale_ssr, double beta0, uint64_t key,
                          uint64_t iteration, uint32_t purpose, int* done) {
  if (!e || !done) return fail("null argument");
  *done = 0;
  static const bool off = std::getenv("BFMMM_NO_SIGMA_TAIL") != nullptr;
  if (off || !e->mom_enabled || (need_exchange && !e->xchg_on)) return 0;
  CU(cudaSetDevice(e->device));
  bf::SigmaTail t;
  std::memset(&t, 0, sizeof(t));
  t.on = 1; t.shape = a_shape; t.scale_ssr = scale_ssr; t.beta0 = beta0; t.key = key; t.iteration = iteration; t.purpose = purpose;
  e->sig_seq += 1.0;
  t.sigma_dev = e->sigma_dev; t.host = e->h_sig_dev; t.seq = e->sig_seq; t.world = 1;
  if (need_exchange) {
    t.peers = e->xchg_peers; t.rank = e->xchg_rank; t.world = e->xchg_world; t.cap = e->xchg_cap; t.xseq = ++*e->xchg_seq;
  }
  if (ssr_launch(e, &t)) return 1;
  e->sigma_armed = true;
  *done = 1;
  return 0;
}
int64_t bfmmm_stats_len(bfmmm_engine* e) { return e ? e->stats_len : 0; }
// 1 when the statistics pass queued last already summed the whole buffer over the shards in its epilogue
int bfmmm_stats_exchanged(bfmmm_engine* e) { return e && e->stats_exchanged ? 1 : 0; }
// Fuses the peer-memory exchange (p2p_hook.cu's mailboxes and sequence counter) into the statistics pass's epilogue.
int bfmmm_engine_set_exchange(bfmmm_engine* e, const bf::P2PPeers* peers, int rank, int world, int cap, unsigned long long* seq) {
  if (!e) return fail("null engine");
  if (!peers || world < 2) { e->xchg_on = false; return 0; }
  if (world > bf::P2P_MAX_RANKS || rank < 0 || rank >= world || !seq) return fail("bfmmm_engine_set_exchange: bad argument");
  if (e->ragged) { e->xchg_on = false; return 0; }
  if (cap < e->K + 3 + e->q * e->q + e->P * e->q) return fail("bfmmm_engine_set_exchange: mailbox slots shorter than the statistics buffer");
  e->xchg_peers = *peers; e->xchg_rank = rank; e->xchg_world = world; e->xchg_cap = cap; e->xchg_seq = seq;
  e->xchg_on = !std::getenv("BFMMM_NO_FUSED_EXCHANGE");
  return 0;
}
int bfmmm_engine_devinfo(bfmmm_engine* e, bf::EngineDevInfo* o) {
  if (!e || !o) return fail("null argument");
  o->stats = e->stats; o->glob = e->glob; o->Pc = e->Pc; o->P4 = (e->Pc + 3) & ~3; o->QS = e->QS; o->hbL = e->hbL;
  o->L_host = e->L.data(); o->stream = e->stream; o->device = e->device; o->stats_len = e->stats_len;
  return 0;
}
// The proposal half of the NEXT Z step, on the side stream: it starts once the last writer of Z (the Z step in
// flight) has finished and runs beside whatever the main stream does next.  (pi, alpha_3, a_Z_PM, iteration) must be the ones the step will be
// asked for with; otherwise bfmmm_update_z* makes its own proposal and this one is dropped.
int bfmmm_z_propose_async(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, uint64_t iteration, bool after_queued) {
  if (!e || !pi) return fail("null argument");
  if (!e->z_split) return 0;
  CU(cudaSetDevice(e->device));
  if (z_split_init(e)) return 1;
  bf::PassArgs a;
  fill_pass(e, a, 1.0);
  a.iteration = iteration;
  z_fill(e, a, pi, alpha3, a_Z_PM, nullptr);
  a.zprop_out = e->zprop;
  CU(cudaStreamWaitEvent(e->side, e->ev_zdone, 0));                // the last writer of Z (z_written)
  if (after_queued) {
    // ... and whatever the engine's stream holds right now (the statistics kernel): it shares the FP64 pipe with this
    // kernel, so running beside it only stretches the critical path; the gap that follows it is free
    CU(cudaEventRecord(e->ev_queued, e->stream));
    CU(cudaStreamWaitEvent(e->side, e->ev_queued, 0));
  }
  int rc = bf::launch_z_propose(a, e->K, e->side);
  if (rc) return fail("z proposal kernel launch failed rc=" + std::to_string(rc));
  CU(cudaEventRecord(e->ev_prop, e->side));
  for (int k = 0; k < e->K; k++) e->prop_pi[k] = pi[k];
  e->prop_alpha3 = alpha3; e->prop_a = a_Z_PM; e->prop_key = e->key; e->prop_iter = iteration;
  e->prop_valid = true;
  return 0;
}

// [sum_i log Z_ik (K) | accepts] of the Z step just queued, copied to the host right behind it (the statistics kernel
// queued next does not delay it); _wait blocks until they have arrived.  Single-shard chains only: with several shards
// the sums must go through the exchange first.
int bfmmm_slz_read_begin(bfmmm_engine* e) {
  if (!e) return fail("null engine");
  dbg_stale("slz_read_begin entry");
  CU(cudaSetDevice(e->device));
  if (z_split_init(e)) return 1;
  if (bf::launch_copy_to_host(e->stats + e->off_slz(), e->h_slz_dev, e->K + 1, e->stream)) return fail("copy_to_host kernel launch failed");
  CU(cudaEventRecord(e->ev_slz, e->stream));
  dbg_stale("slz_read_begin exit");
  return 0;
}
int bfmmm_slz_read_wait(bfmmm_engine* e, double* out) {
  if (!e || !e->ev_slz) return fail("bfmmm_slz_read_wait: nothing to wait for");
  CU(cudaSetDevice(e->device));
  CU(cudaEventSynchronize(e->ev_slz));
  std::copy(e->h_slz, e->h_slz + e->K + 1, out);
  return 0;
}
bool bfmmm_z_ahead_supported(bfmmm_engine* e) { return e && e->z_split; }
// duration in microseconds of the last proposal kernel that ran on the engine's own stream (alone, in front of its accept
// kernel), or a negative value when none has finished yet: what the sampler weighs against its own block-draw time
double bfmmm_z_propose_us(bfmmm_engine* e) {
  if (!e || !e->k1_timed) return -1.0;
  cudaSetDevice(e->device);
  float ms = 0;
  cudaError_t rc = cudaEventQuery(e->ev_k1b);
  if (rc == cudaSuccess) rc = cudaEventElapsedTime(&ms, e->ev_k1a, e->ev_k1b);
  if (rc != cudaSuccess) {
    if (rc != cudaErrorNotReady && std::getenv("BFMMM_DEBUG")) std::fprintf(stderr, "bfmmm_z_propose_us: %s\n", cudaGetErrorString(rc));
    cudaGetLastError();          // not an error of the sweep: do not leave it for the next launch check to find
    return -1.0;
  }
  return 1e3 * ms;
}
void bfmmm_moments_invalidate(bfmmm_engine* e) { if (e) e->mom_valid = false; }
// Z step with pi, alpha_3 and sigma^2 read from device memory ([pi (8) | alpha_3 | sigma^2])
int bfmmm_update_z_async_p(bfmmm_engine* e, double a_Z_PM, double beta, const double* zpar_dev) {
  if (!e || !zpar_dev) return fail("null argument");
  CU(cudaSetDevice(e->device));
  return z_launch(e, nullptr, 1.0, a_Z_PM, beta, false, false, zpar_dev);
}
// chi sweep with sigma^2 read from device memory
int bfmmm_update_chi_async_p(bfmmm_engine* e, double beta, const double* sigma_dev) {
  if (!e || !sigma_dev) return fail("null argument");
  CU(cudaSetDevice(e->device));
  return chi_launch(e, beta, false, sigma_dev);
}
// the value of sigma^2 the engine passes by value (marginal log-likelihood, Z step through the C ABI)
int bfmmm_set_sigma(bfmmm_engine* e, double sigma_sq) {
  if (!e || !(sigma_sq > 0)) return fail("bfmmm_set_sigma: bad argument");
  e->sigma_sq = sigma_sq;
  return 0;
}

namespace bf {
int pass_grid(int ld, int v) { return (ld + PF_THREADS * v - 1) / (PF_THREADS * v); }
}
