// stats_kernels.cu -- sufficient-statistic contraction on the FP64 tensor-core path (DMMA).
//
// Feeds updateNu / updatePhi / updateEta / updateXi (reference UpdateNu.h:42-63,
// UpdatePhi.h:44-71, UpdateEta.h:51-81, UpdateXi.h:51-72).  Those loops accumulate, per block a,
//   M_a = sum_i w_ia^2 B_i'B_i,   m_a = sum_i w_ia B_i'(y_i - B_i sum_{b != a} w_ib c_b)
// with one full data pass per block.  On a common basis all of them are functions of
//   W'W  (q x q)   and   C~'W  (P x q),     W[i][f] = feature weight, C~ = whitened coefficients,
// which this kernel produces in ONE pass:  [C~ ; W]' W  is a GEMM whose K dimension is the
// function index (split-K over the whole grid).
//
// Mapping onto mma.sync.m8n8k4.f64 (SASS DMMA): the function index is the MMA K dimension, and
// because a sum over functions is order-free the four K slots of one MMA are fed with functions
// {2c} (and the next MMA with {2c+1}) of an 8-function chunk, where c = lane%4.  Every thread
// therefore issues only 16-byte loads, four lanes cover 64 contiguous bytes of a row, and no
// shared-memory transpose is needed:
//   A fragment (8x4)  a = C~[p = 8*mt + lane/4][function slot c]         (rows p >= P are zero)
//   B fragment (4x8)  b = W [function slot c][feature f = 8*nt + lane/4] (computed on the fly
//                         from Z, chi, X: w = Z_k * chi_m * x_d)
//   W'W tiles reuse the same registers: A = W' fragment (row = feature, col = slot) == b.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace bf {

static inline void stats_shape(int P, int q, int& MT, int& NT, int& gy);

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Operands travel global -> shared with cp.async (LDGSTS, 16 bytes per thread, L1 bypassed) into a
// ring of ST stages of THREAD-PRIVATE slots: nothing is held in registers while in flight, so two
// blocks of 8 warps are resident per SM with ST-1 chunks per warp in flight (the first version kept
// the ring in registers: 170 registers, 8 warps/SM, long-scoreboard 55 %).  A thread only ever reads
// the slots it filled itself, so no block barrier is needed: cp.async.wait_group orders its own copies.
template <int MT, int NT, int ST, bool HASX>
__global__ void __launch_bounds__(ST_THREADS, 2) stats_kernel(const StatsArgs a) {
  constexpr int TILES = MT * NT + NT * NT;
  constexpr int PER = HASX ? 3 : 2;            // slots per n-tile: Z, chi (, x)
  constexpr int ITEMS = MT + PER * NT;
  extern __shared__ double2 ring[];            // [ST][ITEMS][ST_THREADS]
  __shared__ double s_acc[TILES * 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int mt0 = blockIdx.y * MT;
  const bool do_wtw = (blockIdx.y == 0);

  // feature owned by this thread in each n-tile
  const double* zp[NT];
  const double* cp[NT];
  const double* xp[NT];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    int f = nt * 8 + g;
    zp[nt] = nullptr; cp[nt] = nullptr; xp[nt] = nullptr;
    if (f < a.q) {
      int dd = f % (1 + a.D);
      int km = f / (1 + a.D);
      int mm = km % (a.M + 1), k = km / (a.M + 1);
      zp[nt] = a.Z + (size_t)k * a.ld;
      if (mm > 0) cp[nt] = a.chi + (size_t)(mm - 1) * a.ld;
      if (dd > 0) xp[nt] = a.X + (size_t)(dd - 1) * a.ld;
    }
  }
  const double* ap[MT];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    int p = (mt0 + mt) * 8 + g;
    ap[mt] = (p < a.P) ? a.Ct + (size_t)p * a.ld : nullptr;
  }

  double R[MT][NT][2], S[NT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { R[mt][nt][0] = 0; R[mt][nt][1] = 0; }
#pragma unroll
  for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
    for (int n2 = 0; n2 < NT; n2++) { S[n1][n2][0] = 0; S[n1][n2][1] = 0; }

  const int n_chunks = a.ld >> 3;
  const int wstride = gridDim.x * ST_WARPS;
  auto slot = [&](int stage, int item) { return ring + ((size_t)(stage * ITEMS + item) * ST_THREADS + threadIdx.x); };
  auto issue = [&](int stage, int chunk) {
    if (chunk < n_chunks) {
      const int i = (chunk << 3) + 2 * c;
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
        if (ap[mt]) cp_async16(slot(stage, mt), ap[mt] + i);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        if (zp[nt]) cp_async16(slot(stage, MT + PER * nt), zp[nt] + i);
        if (cp[nt]) cp_async16(slot(stage, MT + PER * nt + 1), cp[nt] + i);
        if (HASX && xp[nt]) cp_async16(slot(stage, MT + PER * nt + 2), xp[nt] + i);
      }
    }
    cp_async_commit();          // one group per call (possibly empty) keeps the group count uniform
  };
  int ch = blockIdx.x * ST_WARPS + warp;
#pragma unroll
  for (int st = 0; st < ST - 1; st++) issue(st, ch + st * wstride);
  int stage = 0;
  for (; ch < n_chunks; ch += wstride) {
    int nxt = stage + (ST - 1); if (nxt >= ST) nxt -= ST;
    issue(nxt, ch + (ST - 1) * wstride);
    cp_async_wait<ST - 1>();    // the group of the current chunk has landed
    double2 av[MT], wv[NT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) av[mt] = ap[mt] ? *slot(stage, mt) : make_double2(0.0, 0.0);
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      double2 w = make_double2(0.0, 0.0);
      if (zp[nt]) {
        w = *slot(stage, MT + PER * nt);
        if (cp[nt]) { const double2 t = *slot(stage, MT + PER * nt + 1); w.x *= t.x; w.y *= t.y; }
        if (HASX && xp[nt]) { const double2 t = *slot(stage, MT + PER * nt + 2); w.x *= t.x; w.y *= t.y; }
      }
      wv[nt] = w;
    }
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        dmma884(R[mt][nt][0], R[mt][nt][1], av[mt].x, wv[nt].x);
        dmma884(R[mt][nt][0], R[mt][nt][1], av[mt].y, wv[nt].y);
      }
    if (do_wtw) {
#pragma unroll
      for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
        for (int n2 = n1; n2 < NT; n2++) {
          dmma884(S[n1][n2][0], S[n1][n2][1], wv[n1].x, wv[n2].x);
          dmma884(S[n1][n2][0], S[n1][n2][1], wv[n1].y, wv[n2].y);
        }
    }
    stage = stage + 1 == ST ? 0 : stage + 1;
  }
  cp_async_wait<0>();

  // block reduction: warps add their fragments into shared memory one after the other (fixed order)
  for (int w = 0; w < ST_WARPS; w++) {
    if (warp == w) {
      int t = 0;
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++, t++) {
          int idx = t * 64 + g * 8 + 2 * c;
          if (w == 0) { s_acc[idx] = R[mt][nt][0]; s_acc[idx + 1] = R[mt][nt][1]; }
          else { s_acc[idx] += R[mt][nt][0]; s_acc[idx + 1] += R[mt][nt][1]; }
        }
#pragma unroll
      for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
        for (int n2 = 0; n2 < NT; n2++, t++) {
          int idx = t * 64 + g * 8 + 2 * c;
          if (w == 0) { s_acc[idx] = S[n1][n2][0]; s_acc[idx + 1] = S[n1][n2][1]; }
          else { s_acc[idx] += S[n1][n2][0]; s_acc[idx + 1] += S[n1][n2][1]; }
        }
    }
    __syncthreads();
  }
  double* row = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (TILES * 64);
  for (int idx = threadIdx.x; idx < TILES * 64; idx += ST_THREADS) row[idx] = s_acc[idx];
}

// ------------------------------------------------------------------ TMA-fed variant
// The same contraction with the operands staged by the TMA engine.  A producer warp issues, per stage
// of 64 functions, ONE cp.async.bulk.tensor.2d (SASS UTMALDG) per operand -- the [8 MT rows x 72
// functions] box of the coefficient cache, the K rows of Z, the M rows of chi, the D rows of X -- into a
// ring of shared-memory stages and the TMA unit signals the stage's "full" mbarrier with the byte count;
// the eight consumer warps wait on it, read their DMMA fragments from shared memory and release the
// stage through the "empty" mbarrier.  No thread holds operands in registers while they are in flight
// and the loads need no LSU issue slots.  The box is 72 functions wide although a stage consumes 64: the
// 576-byte row pitch puts the two rows a quarter-warp touches into disjoint bank halves (a 512-byte
// pitch would be a 2-way conflict on every fragment load; the 128-byte swizzle needs rows <= 128 B); the
// 8 extra functions are the next stage's first (an L2 hit then).  Rows beyond P and functions beyond ld
// are zero-filled by the TMA unit, so there are no bounds predicates.  (A first version issued one 1-D
// bulk copy per ROW, 26 per stage: a single thread cannot issue them fast enough -- 84 us.)
constexpr int TM_CONSUMERS = 8;                 // consumer warps
constexpr int TM_THREADS = (TM_CONSUMERS + 1) * 32;
constexpr int TM_FN = 64;                       // functions consumed per stage (8 per consumer warp)
constexpr int TM_BOX = 72;                      // functions copied per stage row
constexpr int TM_STRIDE = TM_BOX * 8;           // shared-memory row pitch in bytes (576)
#ifndef BF_TM_STAGES
#define BF_TM_STAGES 4
#endif
constexpr int TM_STAGES = BF_TM_STAGES;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  for (unsigned spin = 0; spin < (1u << 28); spin++) {
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__host__ __device__ inline int tm_align128(int bytes) { return (bytes + 127) & ~127; }

// FOLD (all cache tiles in one block, P % 8 and q % 8 both in 1..4): the last cache tile has at most four real rows and
// the last feature tile at most four real features, so the remnant features ride in rows 4..7 of the last cache tile's
// A fragment.  The tiles (MT-1, nt) then carry W'W[remnant][tile nt] in those rows and the NT tiles (n1, NT-1) of W'W
// are not issued: 7 instead of 9 tiles (14 instead of 18 DMMAs per 8 functions) at P = 20, q = 12.
template <int MT, int NT, bool FOLD>
#ifndef BF_ST_BPS
#define BF_ST_BPS 2      // resident blocks per SM of the statistics kernels (grid = BF_ST_BPS * SM count)
#endif
__global__ void __launch_bounds__(TM_THREADS, BF_ST_BPS) stats_kernel_tma(const StatsArgs a, const __grid_constant__ StatsTmaMaps tm) {
  constexpr int TILES = MT * NT + NT * NT;
  extern __shared__ __align__(128) unsigned char tm_smem[];
  __shared__ double s_acc[TILES * 64];
  __shared__ unsigned long long full_bar[TM_STAGES], empty_bar[TM_STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int mt0 = blockIdx.y * MT;
  const bool do_wtw = (blockIdx.y == 0);
  // regions of one stage (each 128-byte aligned): MT*8 rows of the cache, then Z (K), chi (M), X (D)
  const int off_z = tm_align128(MT * 8 * TM_STRIDE);
  const int off_c = off_z + tm_align128(a.K * TM_STRIDE);
  const int off_x = off_c + tm_align128(a.M * TM_STRIDE);
  const int stage_bytes = off_x + tm_align128(a.D * TM_STRIDE);
  const unsigned tx_bytes = (unsigned)((MT * 8 + a.K + a.M + a.D) * TM_STRIDE);
  if (threadIdx.x == 0) {
    for (int s = 0; s < TM_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TM_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_super = a.ld / TM_FN;                 // ld is a multiple of 64
  if (warp == TM_CONSUMERS) {
    // ===== producer warp: one lane drives the TMA engine =====
    if (lane == 0) {
      int it = 0;
      for (int sc = blockIdx.x; sc < n_super; sc += gridDim.x, it++) {
        const int st = it % TM_STAGES;
        if (it >= TM_STAGES) mbar_wait(&empty_bar[st], ((it / TM_STAGES) - 1) & 1);
        mbar_expect_tx(&full_bar[st], tx_bytes);
        unsigned char* base = tm_smem + (size_t)st * stage_bytes;
        const int col = sc * TM_FN;
        tma_load_2d(base, &tm.ct, col, mt0 * 8, &full_bar[st]);
        tma_load_2d(base + off_z, &tm.z, col, 0, &full_bar[st]);
        tma_load_2d(base + off_c, &tm.chi, col, 0, &full_bar[st]);
        if (a.D > 0) tma_load_2d(base + off_x, &tm.x, col, 0, &full_bar[st]);
      }
    }
  } else {
    // ===== consumer warps =====
    int zoff[NT], coff_[NT], xoff[NT];       // byte offsets of this thread's feature rows in a stage (-1: none)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      int f = nt * 8 + g;
      zoff[nt] = -1; coff_[nt] = -1; xoff[nt] = -1;
      if (f < a.q) {
        int dd = f % (1 + a.D), km = f / (1 + a.D), mm = km % (a.M + 1), k = km / (a.M + 1);
        zoff[nt] = off_z + k * TM_STRIDE;
        if (mm > 0) coff_[nt] = off_c + (mm - 1) * TM_STRIDE;
        if (dd > 0) xoff[nt] = off_x + (dd - 1) * TM_STRIDE;
      }
    }
    int zoffF = -1, coffF = -1, xoffF = -1;  // FOLD: the remnant feature this lane (g >= 4) carries in the last cache tile
    if (FOLD && g >= 4) {
      int f = (NT - 1) * 8 + (g - 4);
      if (f < a.q) {
        int dd = f % (1 + a.D), km = f / (1 + a.D), mm = km % (a.M + 1), k = km / (a.M + 1);
        zoffF = off_z + k * TM_STRIDE;
        if (mm > 0) coffF = off_c + (mm - 1) * TM_STRIDE;
        if (dd > 0) xoffF = off_x + (dd - 1) * TM_STRIDE;
      }
    }
    double R[MT][NT][2], S[NT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) { R[mt][nt][0] = 0; R[mt][nt][1] = 0; }
#pragma unroll
    for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
      for (int n2 = 0; n2 < NT; n2++) { S[n1][n2][0] = 0; S[n1][n2][1] = 0; }
    const int coff = warp * 64 + c * 16;            // byte offset of this thread's two functions in a row
    int it = 0;
    for (int sc = blockIdx.x; sc < n_super; sc += gridDim.x, it++) {
      const int st = it % TM_STAGES;
      mbar_wait(&full_bar[st], (it / TM_STAGES) & 1);
      const unsigned char* base = tm_smem + (size_t)st * stage_bytes + coff;
      double2 av[MT], wv[NT];
#pragma unroll
      for (int mt = 0; mt < MT; mt++)                 // rows >= P were zero-filled by the TMA unit
        av[mt] = *reinterpret_cast<const double2*>(base + (size_t)(mt * 8 + g) * TM_STRIDE);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        double2 w = make_double2(0.0, 0.0);
        if (zoff[nt] >= 0) {
          w = *reinterpret_cast<const double2*>(base + zoff[nt]);
          if (coff_[nt] >= 0) { const double2 t = *reinterpret_cast<const double2*>(base + coff_[nt]); w.x *= t.x; w.y *= t.y; }
          if (xoff[nt] >= 0) { const double2 t = *reinterpret_cast<const double2*>(base + xoff[nt]); w.x *= t.x; w.y *= t.y; }
        }
        wv[nt] = w;
      }
      if (FOLD && zoffF >= 0) {                       // rows 4..7 of the last cache tile (zero rows >= P) become W' rows
        double2 w = *reinterpret_cast<const double2*>(base + zoffF);
        if (coffF >= 0) { const double2 t = *reinterpret_cast<const double2*>(base + coffF); w.x *= t.x; w.y *= t.y; }
        if (xoffF >= 0) { const double2 t = *reinterpret_cast<const double2*>(base + xoffF); w.x *= t.x; w.y *= t.y; }
        av[MT - 1] = w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);   // operands are in registers: the stage may be refilled
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          dmma884(R[mt][nt][0], R[mt][nt][1], av[mt].x, wv[nt].x);
          dmma884(R[mt][nt][0], R[mt][nt][1], av[mt].y, wv[nt].y);
        }
      if (do_wtw) {
#pragma unroll
        for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
          for (int n2 = n1; n2 < (FOLD ? NT - 1 : NT); n2++) {
            dmma884(S[n1][n2][0], S[n1][n2][1], wv[n1].x, wv[n2].x);
            dmma884(S[n1][n2][0], S[n1][n2][1], wv[n1].y, wv[n2].y);
          }
      }
    }
    // block reduction over the consumer warps (fixed order); the producer warp joins the barriers
    for (int w = 0; w < TM_CONSUMERS; w++) {
      if (warp == w) {
        int t = 0;
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
          for (int nt = 0; nt < NT; nt++, t++) {
            int idx = t * 64 + g * 8 + 2 * c;
            if (w == 0) { s_acc[idx] = R[mt][nt][0]; s_acc[idx + 1] = R[mt][nt][1]; }
            else { s_acc[idx] += R[mt][nt][0]; s_acc[idx + 1] += R[mt][nt][1]; }
          }
#pragma unroll
        for (int n1 = 0; n1 < NT; n1++)
#pragma unroll
          for (int n2 = 0; n2 < NT; n2++, t++) {
            int idx = t * 64 + g * 8 + 2 * c;
            if (w == 0) { s_acc[idx] = S[n1][n2][0]; s_acc[idx + 1] = S[n1][n2][1]; }
            else { s_acc[idx] += S[n1][n2][0]; s_acc[idx + 1] += S[n1][n2][1]; }
          }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TM_CONSUMERS * 32) : "memory");   // consumers only
    }
  }
  __syncthreads();
  double* row = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (TILES * 64);
  for (int idx = threadIdx.x; idx < TILES * 64; idx += TM_THREADS) row[idx] = s_acc[idx];
}

// second stage: one warp per output element sums the block partials (lane-strided, then a shuffle
// tree: a fixed order for a fixed grid, so the result is reproducible).
//
// Epilogue (a.ep): the sampler reads the reduced buffer on the host right behind this kernel, and on several GPUs sums
// it over the shards first.  Both used to be kernels of their own (p2p_allreduce_kernel, copy_to_host_kernel); here the
// last block to finish (ticket) does their work:
//   * one GPU: it copies header and statistics into the mapped host copy (coalesced; element-wise 8-byte stores from
//     the 48 reducing blocks were measured slower than the separate copy kernel: +4 us);
//   * several GPUs: it stores the buffer into slot [rank] of EVERY rank's mailbox (remote stores over NVLink), publishes
//     this rank's sequence number, waits for the others', sums the slots in rank order -- bit-identical on all ranks --
//     and writes the totals to `stats` and to the host copy.
// One launch instead of three between the contraction and the host's block draws.
template <int MT, int NT>
__global__ void __launch_bounds__(256) stats_final_kernel(const StatsArgs a, int gx, int fold) {
  constexpr int TILES = MT * NT + NT * NT;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nR = a.P * a.q, nS = a.q * a.q;
  if (e < nR + nS) {
    int by, idx;
    if (e < nR) {
      int p = e % a.P, f = e / a.P;
      int mtg = p >> 3;
      by = mtg / MT;
      idx = ((mtg % MT) * NT + (f >> 3)) * 64 + (p & 7) * 8 + (f & 7);
    } else {
      int s = e - nR;
      int f1 = s % a.q, f2 = s / a.q;
      if (f1 > f2) { int t = f1; f1 = f2; f2 = t; }     // only tiles n1 <= n2 are accumulated
      by = 0;
      idx = (MT * NT + (f1 >> 3) * NT + (f2 >> 3)) * 64 + (f1 & 7) * 8 + (f2 & 7);
      // folded layout: W'W[f1][f2] with f2 in the remnant tile sits in row 4 + (f2 & 7) of cache tile (MT-1, f1 >> 3)
      if (fold && (f2 >> 3) == NT - 1) idx = ((MT - 1) * NT + (f1 >> 3)) * 64 + (4 + (f2 & 7)) * 8 + (f1 & 7);
    }
    const double* base = a.partials + (size_t)by * gx * (TILES * 64) + idx;
    double t = 0;
    for (int b = lane; b < gx; b += 32) t += base[(size_t)b * (TILES * 64)];
    t = warp_sum(t);
    if (lane == 0) { if (e < nR) a.CtW[e] = t; else a.WtW[e - nR] = t; }
  }
  const StatsEpilogue& ep = a.ep;
  const bool xchg = ep.world > 1;
  if (!xchg && !ep.mirror) return;
  // ---- epilogue: the last block to arrive sees every block's elements (fence + ticket)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ep.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) *ep.ticket = 0;
  const int len = ep.hdr + nS + nR;          // [header (written by the Z / chi / SSR kernels before this one) | W'W | C~'W]
  if (!xchg) {
    for (int i = threadIdx.x; i < len; i += blockDim.x) ep.mirror[i] = __ldcg(ep.stats + i);
    return;
  }
  const int par = (int)(ep.seq & 1ull);
  for (int r = 0; r < ep.world; r++) {       // (a) my partial sums into slot [par][rank] of every mailbox
    double* dst = reinterpret_cast<double*>(ep.peers.box[r] + P2P_HDR) + ((size_t)par * ep.world + ep.rank) * ep.cap;
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldcg(ep.stats + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < ep.world) {              // (b) publish, (c) wait (bounded: a missing rank traps instead of hanging the device)
    volatile unsigned long long* theirs = reinterpret_cast<volatile unsigned long long*>(ep.peers.box[threadIdx.x]) + ep.rank;
    *theirs = ep.seq;
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(ep.peers.box[ep.rank]) + threadIdx.x;
    unsigned long long spins = 0;
    while (*mine < ep.seq) {
      if (++spins > (1ull << 31)) __trap();
    }
  }
  __threadfence_system();
  __syncthreads();
  // (d) sum in rank order
  const double* src = reinterpret_cast<const double*>(ep.peers.box[ep.rank] + P2P_HDR) + (size_t)par * ep.world * ep.cap;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    double t = 0;
    for (int r = 0; r < ep.world; r++) t += __ldcg(src + (size_t)r * ep.cap + i);
    ep.stats[i] = t;
    if (ep.mirror) ep.mirror[i] = t;
  }
}

int stats_blocks(int sm_count) { return BF_ST_BPS * sm_count; }

static inline void stats_shape(int P, int q, int& MT, int& NT, int& gy) {
  NT = (q + 7) / 8;
  int mtiles = (P + 7) / 8;
  MT = mtiles < 4 ? mtiles : 4;
  if (NT >= 4 && MT > 2) MT = 2;     // keep the accumulator file within the register budget
  gy = (mtiles + MT - 1) / MT;
}

size_t stats_partial_doubles(int P, int q, int blocks) {
  int MT, NT, gy;
  stats_shape(P, q, MT, NT, gy);
  return (size_t)gy * blocks * (MT * NT + NT * NT) * 64;
}

template <int MT, int NT, bool HASX>
static int launch_stats_x(const StatsArgs& a, int gy, cudaStream_t s) {
  dim3 grid(a.blocks, gy);
  // three ring stages when two blocks of them fit in one SM's shared memory, otherwise two
  constexpr size_t stage_bytes = (size_t)(MT + (HASX ? 3 : 2) * NT) * ST_THREADS * sizeof(double2);
  constexpr bool three = 3 * stage_bytes <= 104 * 1024;
  constexpr int ST = three ? 3 : 2;
  const size_t smem = ST * stage_bytes;
  int dev = 0;
  cudaGetDevice(&dev);
  static std::set<int> configured;                            // per device: the attribute is per context
  static std::mutex configured_mu;                            // (several host threads may drive several engines)
  {
    std::lock_guard<std::mutex> lock(configured_mu);
    if (!configured.count(dev)) {
      cudaError_t e = cudaFuncSetAttribute(stats_kernel<MT, NT, ST, HASX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      configured.insert(dev);
    }
  }
  stats_kernel<MT, NT, ST, HASX><<<grid, ST_THREADS, smem, s>>>(a);
  return 0;
}

template <int MT, int NT>
static int launch_stats_tma(const StatsArgs& a, int gy, cudaStream_t s, bool& used, bool& folded) {
  used = false; folded = false;
  static const bool disabled = std::getenv("BFMMM_STATS_NO_TMA") != nullptr;
  if (disabled || !a.tma || !a.tma->valid || a.tma->mt != MT || (a.ld % TM_FN) != 0) return 0;
  const size_t stage = (size_t)tm_align128(MT * 8 * TM_STRIDE) + tm_align128(a.K * TM_STRIDE) + tm_align128(a.M * TM_STRIDE) +
                       tm_align128(a.D * TM_STRIDE);
  const size_t smem = (size_t)TM_STAGES * stage;
  if (smem > (BF_ST_BPS == 2 ? 100 : 64) * 1024) return 0;    // BF_ST_BPS blocks per SM
  int dev = 0;
  cudaGetDevice(&dev);
  // The attribute is one value per (kernel, device) and the stage size depends on K, M, D, not only on <MT, NT>: it is
  // raised when an engine needs more than any before it (a set of the sizes seen would let a smaller engine lower it
  // under a larger one created earlier).  Several host threads may drive several engines: locked.
  static const bool no_fold = std::getenv("BFMMM_STATS_NO_FOLD") != nullptr;
  const int pr = a.P & 7, qr = a.q & 7;
  folded = !no_fold && gy == 1 && pr >= 1 && pr <= 4 && qr >= 1 && qr <= 4;
  static std::map<std::pair<int, bool>, size_t> granted;
  static std::mutex granted_mu;
  {
    std::lock_guard<std::mutex> lock(granted_mu);
    auto it = granted.find({dev, folded});
    if (it == granted.end() || it->second < smem) {
      cudaError_t e = folded ? cudaFuncSetAttribute(stats_kernel_tma<MT, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(stats_kernel_tma<MT, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      granted[{dev, folded}] = smem;
    }
  }
  dim3 grid(a.blocks, gy);
  if (folded) stats_kernel_tma<MT, NT, true><<<grid, TM_THREADS, smem, s>>>(a, *a.tma);
  else stats_kernel_tma<MT, NT, false><<<grid, TM_THREADS, smem, s>>>(a, *a.tma);
  used = true;
  return 0;
}

template <int MT, int NT>
static int launch_stats_t(const StatsArgs& a, int gy, cudaStream_t s) {
  bool used = false, folded = false;
  int rc = launch_stats_tma<MT, NT>(a, gy, s, used, folded);
  if (rc) return rc;
  if (!used) rc = a.D > 0 ? launch_stats_x<MT, NT, true>(a, gy, s) : launch_stats_x<MT, NT, false>(a, gy, s);
  if (rc) return rc;
  int tot = a.P * a.q + a.q * a.q;
  stats_final_kernel<MT, NT><<<(tot + 7) / 8, 256, 0, s>>>(a, a.blocks, folded ? 1 : 0);
  g_launch_count += 2;
  return (int)cudaGetLastError();
}

// Tensor maps of the four operands ([rows][ld] doubles, function index contiguous): box = 72 functions x
// (8 MT | K | M | D) rows.  Encoded once per engine through the driver entry point (no libcuda link).
int stats_tma_setup(StatsTmaMaps* out, const double* Ct, const double* Z, const double* chi, const double* X,
                    int ld, int P, int K, int M, int D, int q) {
  out->valid = 0;
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return 1;
  int MT, NT, gy;
  stats_shape(P, q, MT, NT, gy);
  auto enc = [&](CUtensorMap* m, const double* base, int rows, int box_rows) -> bool {
    if (rows <= 0) return true;
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)TM_BOX, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return ((encode_fn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!enc(&out->ct, Ct, P, MT * 8) || !enc(&out->z, Z, K, K) || !enc(&out->chi, chi, M, M) || !enc(&out->x, X, D, D)) return 1;
  out->mt = MT;
  out->valid = 1;
  return 0;
}

int launch_stats(const StatsArgs& a, cudaStream_t s) {
  int MT, NT, gy;
  stats_shape(a.P, a.q, MT, NT, gy);
#define BF_ST(M_, N_) if (MT == M_ && NT == N_) return launch_stats_t<M_, N_>(a, gy, s);
  BF_ST(1, 1) BF_ST(2, 1) BF_ST(3, 1) BF_ST(4, 1)
  BF_ST(1, 2) BF_ST(2, 2) BF_ST(3, 2) BF_ST(4, 2)
  BF_ST(1, 3) BF_ST(2, 3) BF_ST(3, 3) BF_ST(4, 3)
  BF_ST(1, 4) BF_ST(2, 4)
  BF_ST(1, 5) BF_ST(2, 5)
  BF_ST(1, 6) BF_ST(2, 6)
#undef BF_ST
  return -3;
}

}  // namespace bf
