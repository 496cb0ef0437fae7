// common.cuh -- shared definitions of the sm_100a kernels behind include/bfmmm.h.
//
// Data layout in HBM:
//   Ct   [P][ld]  whitened projected coefficients  c~_i = L^{-1} B' y_i   (SoA: coefficient-major,
//                 function index contiguous, ld = n rounded up to 64) -- the per-iteration kernels
//                 stream this cache instead of the raw observations.
//   rss  [ld]     ||y_i - B c_i||^2, the part of the residual orthogonal to the basis
//   Z    [K][ld], chi [M][ld], X [D][ld]   (the reference's column-major n x K etc. with padded ld)
//   glob [P][QS]  whitened global coefficients, feature-minor:
//                 f = ((k*(M+1) + m')*(1+D) + d'), m'=0 mean block / m'=m+1 eigen block m,
//                 d'=0 plain / d'=d+1 covariate d; QS = q rounded up to 2.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace bf {

constexpr int DMAX = 4;         // covariates supported by the covariate-adjusted kernels
constexpr int PF_THREADS = 128; // block size of the per-function passes
constexpr int RED_MAX = 8;      // values reduced per block in the per-function passes (K+1 <= 8)

struct PassArgs {
  int n, ld, P, D, QS;
  int P4;                           // P rounded up to 4: the cache and the globals carry zero rows P..P4-1 (common grid)
  int sm_count, max_blocks;         // launch geometry: persistent grid of at most max_blocks blocks
  int grid_reserve;                 // blocks left out of the resident wave (a side-stream kernel occupies part of one SM)
  const double* __restrict__ Ct;    // common basis: whitened c~ ; ragged grids: least-squares c_i
  const double* __restrict__ Gl;    // ragged grids: lower band of G_i, row (j*P + p) = G_i[p-j][p]; else nullptr
  int bw;                           // ragged grids: band width (degree + 1)
  const double* __restrict__ rss;
  double* __restrict__ Z;
  double* __restrict__ chi;
  const double* __restrict__ X;
  const double* __restrict__ glob;
  double sigma_sq, beta;
  const double* __restrict__ sigma_dev;   // chi: sigma^2 drawn on the device by sigma_draw_kernel (nullptr: use sigma_sq)
  // Z step
  double alpha3, a_Z_PM, log_a_Z_PM, inv_a_Z_PM;
  double c_tot, trigam_a;           // 1 - log a + digamma(a) and trigamma(a) at a = a_Z_PM (host): lgamma(a sum_k z*_k) - lgamma(a sum_k z_k)
                                    // by its Taylor series around a, merged with the Stirling main terms (z_logratio_closed)
  double pi[8];
  const double* __restrict__ zpar_dev;    // Z step: [pi (8) | alpha_3 | sigma^2] in device memory (device-resident sweep) or nullptr
  const double* __restrict__ gam;   // injected draws [K][ld] or nullptr (device RNG)
  const double* __restrict__ u;     // [ld] or nullptr
  const double* __restrict__ eps;   // chi: [M][ld] or nullptr
  const double* __restrict__ zprop; // Z step, accept half: (z*[K] | lr | lu)[ld] made by z_propose_kernel
  double* __restrict__ zprop_out;   // Z step, proposal half: where to leave them
  double* __restrict__ acc_out;     // optional per-function acceptance log-ratio (diagnostics)
  double* __restrict__ draws_out;   // optional: device-RNG draws written back ([K+1][ld] / [M][ld])
  // marginal log-likelihood / CPO accumulation (chi_kernel<..., CPO = true>)
  const double* __restrict__ ni;    // points per function (ragged grids) or nullptr: npts_common
  double npts_common;
  double* __restrict__ logl_out;    // optional per-function marginal log-likelihood
  double* __restrict__ cpo_m;       // running max and scaled sum of exp(-logl) per function (or nullptr)
  double* __restrict__ cpo_s;
  int cpo_first;
  // RNG
  uint64_t key, iteration, global_offset;
  uint32_t rk[20];                  // Philox round keys (k0 + r W0, k1 + r W1), r = 0..9: read as constant operands
  // reduction
  double* __restrict__ partials;    // [gridDim.x][RED_MAX]
  unsigned int* __restrict__ ticket;
  double* __restrict__ out;         // final reduced values (stats buffer slots)
  int n_out;
};

// ------------------------------------------------------------------ Philox4x32-10 (counter-based)
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ static inline void block(uint32_t c[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c[0], hi0, lo0);
      mulhilo(0xCD9E8D57u, c[2], hi1, lo1);
      uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
  }
};

// Stream of doubles for one (function, iteration, purpose): counter = (index lo, index hi,
// iteration*64 + purpose, running block number).  Results depend only on the GLOBAL function
// index, so a chain is independent of how functions are sharded over GPUs.
// Out-of-line FP64 transcendental wrappers.  libdevice inlines ~50-1000 instructions per call of
// log / lgamma / pow; with K- and V-unrolled callers that made the Z kernel 180 KB of SASS and the
// instruction cache (not FP64 or HBM) its bound (profiles/: stall_no_inst).  One copy per kernel.
#ifdef __CUDA_ARCH__
#define BF_NOINLINE __noinline__
#else
#define BF_NOINLINE
#endif
__host__ __device__ BF_NOINLINE inline double nl_log(double x) { return log(x); }

// log(1 + t) for small |t| by its Taylor series (|t| <= 0.05: truncation < 5e-16), else the library log
__host__ __device__ inline double log1p_small(double t) {
  if (fabs(t) > 0.05) return nl_log(1.0 + t);
  double s = -1.0 / 10.0;
  s = fma(s, t, 1.0 / 9.0);  s = fma(s, t, -1.0 / 8.0); s = fma(s, t, 1.0 / 7.0);  s = fma(s, t, -1.0 / 6.0);
  s = fma(s, t, 1.0 / 5.0);  s = fma(s, t, -1.0 / 4.0); s = fma(s, t, 1.0 / 3.0);  s = fma(s, t, -1.0 / 2.0);
  s = fma(s, t, 1.0);
  return s * t;
}

// Counter-based random stream (generic path: host-side global draws, and the rare per-function cases
// the kernels' straight-line fast paths hand over).  One Philox4x32-10 block yields four 32-bit words,
// handed out word by word: a 53-bit uniform takes two words, a 32-bit one a single word.
struct RngStream {
  uint32_t c0, c1, c2, ctr, k0, k1;
  uint32_t w[4];
  int nw;            // words left in w (taken from the top)
  double nspare; int nhave;
  __host__ __device__ RngStream(uint64_t key, uint64_t index, uint64_t iteration, uint32_t purpose)
      : c0((uint32_t)index), c1((uint32_t)(index >> 32)), c2((uint32_t)(iteration * 64 + purpose)), ctr(0),
        k0((uint32_t)key), k1((uint32_t)(key >> 32)), nw(0), nspare(0), nhave(0) { w[0] = w[1] = w[2] = w[3] = 0; }
  __host__ __device__ BF_NOINLINE void refill() {
    uint32_t c[4] = {c0, c1, c2, ctr++};
    Philox::block(c, k0, k1);
    w[0] = c[0]; w[1] = c[1]; w[2] = c[2]; w[3] = c[3];
    nw = 4;
  }
  __host__ __device__ inline uint32_t word() {
    if (nw == 0) refill();
    nw--;
    return nw == 3 ? w[3] : (nw == 2 ? w[2] : (nw == 1 ? w[1] : w[0]));
  }
  __host__ __device__ inline double uniform() {     // (0,1), 53 bits
    uint64_t a = ((uint64_t)word() << 32) | word();
    return ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __host__ __device__ inline double uniform32() {   // (0,1), 32 bits
    return ((double)word() + 0.5) * (1.0 / 4294967296.0);
  }
  __host__ __device__ BF_NOINLINE double normal() {  // Box-Muller; both outputs are used
    if (nhave) { nhave = 0; return nspare; }
    double u1 = uniform(), u2 = uniform32();
    double r = sqrt(-2.0 * log(u1));
    double sn, cs;
#ifdef __CUDA_ARCH__
    sincospi(2.0 * u2, &sn, &cs);
#else
    sn = sin(6.283185307179586476925286766559 * u2); cs = cos(6.283185307179586476925286766559 * u2);
#endif
    nspare = r * sn; nhave = 1;
    return r * cs;
  }
  // Marsaglia-Tsang (2000); shape < 1 boosted by U^(1/shape).  The acceptance test
  //   log U < x^2/2 + d (1 - v + log v) =: R
  // is decided without a logarithm whenever U - 1 < R (accept, since log U <= U - 1) or
  // 1 - 1/U >= R (reject, since log U >= 1 - 1/U); log v = 3 log1p(c x) by its series.
  // If log_out != nullptr it receives log of the returned variate.
  __host__ __device__ BF_NOINLINE double gamma(double shape, double* log_out = nullptr) {
    double boost = 1.0, lboost = 0.0;
    if (shape < 1.0) {
      lboost = nl_log(uniform()) / shape; boost = exp(lboost);
      shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 64; it++) {
      const double x = normal();
      const double t = c * x;
      if (t <= -1.0) continue;
      const double v1 = 1.0 + t;
      const double v = v1 * v1 * v1;
      const double l3 = 3.0 * log1p_small(t);
      const double uu = uniform();
      const double R = 0.5 * x * x + d * (1.0 - v + l3);
      bool ok;
      if (uu - 1.0 < R) ok = true;
      else if (1.0 - 1.0 / uu >= R) ok = false;
      else ok = nl_log(uu) < R;
      if (ok) {
        if (log_out) *log_out = nl_log(d) + l3 + lboost;
        return boost * d * v;
      }
    }
    if (log_out) *log_out = nl_log(boost * d);
    return boost * d;   // unreachable in practice (acceptance > 95% per trial)
  }
};

// ---- straight-line device fast paths (no loops, no calls: every lane of a warp does the same work) ----
#ifdef __CUDACC__
__device__ __forceinline__ void philox_words(uint64_t key, uint64_t index, uint64_t iteration, uint32_t purpose,
                                             uint32_t block, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)(iteration * 64 + purpose), block};
  Philox::block(c, (uint32_t)key, (uint32_t)(key >> 32));
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// The same block with the ten round keys taken from the kernel parameters (host: philox_round_keys): two
// IMAD.WIDE and two LOP3 per round, the key schedule costs nothing per function.
__device__ __forceinline__ void philox_rk(const PassArgs& a, uint64_t index, uint32_t purpose, uint32_t block, uint32_t (&out)[4]) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)(a.iteration * 64 + purpose), c3 = block;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a.rk[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a.rk[2 * r + 1];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  uint64_t a = ((uint64_t)hi << 32) | lo;
  return ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u32(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
#endif
inline void philox_round_keys(uint64_t key, uint32_t (&rk)[20]) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  for (int r = 0; r < 10; r++) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
enum { RNG_Z_PROPOSAL = 1, RNG_Z_ACCEPT = 2, RNG_CHI = 3, RNG_Z_PROPOSAL_SLOW = 4 };

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ld2_stream(const double* p) {      // streamed once: bypass L1
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NV per-thread values: warp shuffle -> shared -> one partial row per
// block -> the last block to finish (ticket) sums the rows in a fixed order and writes `out`.
// returns true in the block that finished last (the one that wrote `out`)
template <int NV>
__device__ __forceinline__ bool grid_reduce_last(double (&v)[NV], const PassArgs& a) {
  __shared__ double s_part[PF_THREADS / 32][RED_MAX];
  __shared__ unsigned int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double w = warp_sum(v[j]);
    if (lane == 0) s_part[warp][j] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w][threadIdx.x];
    a.partials[(size_t)blockIdx.x * RED_MAX + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // last block: NV values, each summed over gridDim.x rows by the whole block in a fixed pattern
  for (int j = 0; j < NV; j++) {
    double t = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      t += __ldcg(&a.partials[(size_t)b * RED_MAX + j]);
    t = warp_sum(t);
    __syncthreads();
    if (lane == 0) s_part[warp][0] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_part[w][0];
      a.out[j] = tot;
    }
  }
  if (threadIdx.x == 0) *a.ticket = 0u;
  return true;
}
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], const PassArgs& a) { (void)grid_reduce_last<NV>(v, a); }

// launchers implemented in the per-kernel translation units
constexpr int BWMAX = 6;        // ragged grids: band width supported by the per-iteration kernels (degree <= 5)
int launch_z_ragged(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_chi_ragged(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_ssr_ragged(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_z(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_z_propose(const PassArgs& a, int K, cudaStream_t s);                          // proposal half (any model)
int launch_z_accept(const PassArgs& a, int K, int M, cudaStream_t s);                    // accept half, common basis
struct SigmaTail;
int launch_moments(const PassArgs& a, int K, int M, double* mom, const SigmaTail* tail, cudaStream_t s);   // moments_kernels.cu
int launch_chi_draw(const PassArgs& a, int K, int M, const double* mom, cudaStream_t s);
int launch_chi(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_ssr(const PassArgs& a, int K, int M, cudaStream_t s);
int launch_mloglik(const PassArgs& a, int K, int M, bool ragged, cudaStream_t s);
int pass_grid(int ld, int v);

// TMA descriptors of the statistics kernel's operands (see stats_kernels.cu)
struct alignas(64) StatsTmaMaps {
  CUtensorMap ct, z, chi, x;
  int valid, mt;
};
// One-shot exchange over NVLink peer memory (p2p_hook.cu): every rank owns a mailbox in its HBM, mapped into its peers.
constexpr int P2P_MAX_RANKS = 16;
constexpr size_t P2P_HDR = 256;     // flags[P2P_MAX_RANKS] (sequence number last published by each rank), then data[2][world][cap]
struct P2PPeers { unsigned char* box[P2P_MAX_RANKS]; };
// what the statistics pass's final-reduction kernel does with the reduced buffer besides leaving it in `stats`
struct StatsEpilogue {
  double* stats;                // device statistics buffer [header (hdr) | W'W | C~'W]
  double* mirror;               // device alias of the mapped page-locked host copy (nullptr: none)
  int hdr;                      // header slots [sum log Z (K) | accepts | ssr | ssr_after]
  unsigned* ticket;             // blocks-done counter (exchange only)
  P2PPeers peers; int rank, world, cap; unsigned long long seq;     // world <= 1: no exchange
};
// updateSigma's draw as the tail of its data pass: the block that finishes the SSR reduction last sums the SSR over the
// shards (peer-memory mailboxes, world > 1), draws sigma^2 (UpdateSigma.h:47-53) and publishes (SSR, sigma^2, seq) in
// mapped host memory -- what p2p_allreduce_kernel + sigma_draw_kernel did as two launches behind the pass.
struct SigmaTail {
  int on;
  double shape, scale_ssr, beta0; uint64_t key, iteration; uint32_t purpose;
  double* sigma_dev; double* host; double seq;
  P2PPeers peers; int rank, world, cap; unsigned long long xseq;
};
struct StatsArgs {
  int n, ld, P, K, M, D, q;
  StatsEpilogue ep;
  const StatsTmaMaps* tma;          // host pointer (passed to the kernel by value) or nullptr
  const double* __restrict__ Ct;
  const double* __restrict__ Z;
  const double* __restrict__ chi;
  const double* __restrict__ X;
  double* __restrict__ partials;   // [blocks][rows]
  double* __restrict__ WtW;        // q x q column-major (device)
  double* __restrict__ CtW;        // P x q column-major (device, whitened)
  int blocks;
};
int launch_stats(const StatsArgs& a, cudaStream_t s);
int stats_blocks(int sm_count);
int stats_tma_setup(StatsTmaMaps* out, const double* Ct, const double* Z, const double* chi, const double* X,
                    int ld, int P, int K, int M, int D, int q);
size_t stats_partial_doubles(int P, int q, int blocks);

struct ProjectArgs {
  int n, ld, P; int64_t T; int64_t i_begin;   // functions [i_begin, i_begin + n) of the shard
  const double* __restrict__ Y;      // n x T row-major (this chunk)
  const double* __restrict__ Q;      // T x P row-major orthonormal basis Q = B L^{-T}
  double* __restrict__ Ct;           // [P][ld]
  double* __restrict__ rss;          // [ld]
};
int launch_project(const ProjectArgs& a, cudaStream_t s);
struct RaggedPrepArgs {
  int n, ld, P, bw, degree, n_knots; int64_t i_begin;
  const int64_t* __restrict__ off;   // n+1 offsets (relative to the y / t / Brows pointers given here)
  const double* __restrict__ y;
  const double* __restrict__ t;      // grid points (spline description) or nullptr
  const double* __restrict__ knots;  // clamped knot vector or nullptr
  const double* __restrict__ Brows;  // user-supplied basis rows (row-major) or nullptr
  double* __restrict__ C;            // [P][ld] least-squares coefficients
  double* __restrict__ H;            // [P][ld] B_i'y_i
  double* __restrict__ Gl;           // [bw*P][ld] lower band of G_i
  double* __restrict__ rss;          // [ld]
};
int launch_ragged_prep(const RaggedPrepArgs& a, cudaStream_t s);
struct RaggedStatsArgs {
  int n, ld, P, bw, K, M, D, q, npairs;
  const double* __restrict__ Gl;
  const double* __restrict__ Z;
  const double* __restrict__ chi;
  const double* __restrict__ X;
  double* __restrict__ partials;
  double* __restrict__ Hb;          // [npairs][bw*P], pair (a <= b) in row-major upper-triangle order
};
int launch_ragged_stats(const RaggedStatsArgs& a, int sm_count, cudaStream_t s);
size_t ragged_stats_partial_doubles(int P, int bw, int q, int sm_count);
int launch_band_width(const double* B, int64_t rows, int P, int* bw_dev, cudaStream_t s);
int launch_bspline(const double* t, int64_t n, const double* knots, int n_knots, int degree, int P,
                   double* B_rowmajor, cudaStream_t s);
int launch_copy_to_host(const double* src, double* dst_mapped, int64_t len, cudaStream_t s);   // SM-driven D2H of a few KB
int launch_sigma_draw(const double* ssr_dev, double a, double scale_ssr, double beta0, uint64_t key, uint64_t iteration,
                      uint32_t purpose, double* sigma_dev, double* host_mapped, double seq, cudaStream_t s);

extern std::atomic<unsigned long long> g_launch_count;   // kernels launched by this library (all engines, all host threads)
int set_error(const char* msg);   // records the message returned by bfmmm_last_error(); returns 1
}  // namespace bf
