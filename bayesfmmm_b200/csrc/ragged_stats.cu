// ragged_stats.cu -- cross-Gram of the Gaussian block updates on RAGGED grids (FP64 DMMA).
//
// On a common grid H_ab = sum_i w_ia w_ib B'B factorises into (W'W)_ab * G (stats_kernels.cu).
// With per-function grids it does not: the reference's sequential block updates (UpdateNu.h:42-63,
// UpdatePhi.h:44-71, UpdateEta.h:51-81, UpdateXi.h:51-72) need, for every pair of features a <= b,
//     H_ab = sum_i w_ia w_ib G_i          (P x P, banded like G_i)
// One pass produces all of them as a GEMM whose K dimension is the function index:
//     Hb[e][pair] = sum_i Gl_i[e] * (w_ia w_ib),     e = j*P + p  <->  G_i[p-j][p]
// A operand = band rows of the cache (the bulk of the bytes, read once per pair-slice),
// B operand = products of feature weights built per 8-function chunk in shared memory.
// Same fragment mapping as stats_kernels.cu (functions {2c}, {2c+1} of a chunk feed the four K slots).
#include "common.cuh"

namespace bf {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MTB = 2;     // m-tiles (16 band rows) per block
constexpr int RS_QMAX = 40;
constexpr int RS_BASEMAX = 16;  // K + M + D <= 6 + 6 + 4

__device__ __forceinline__ void dmma884r(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void rs_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// Work layout: the output (band rows x feature pairs) is cut into S = gy * gz slices of RS_MTB m-tiles x NTB
// pair-tiles (what one warp can accumulate in registers); the EIGHT WARPS of a block take eight different
// slices of the SAME 8-function chunk, so the feature weights of a chunk are formed once per block
// (cooperatively, into double-buffered shared memory) instead of once per slice -- the first version
// recomputed them in every warp: 546 instructions per warp and chunk for 40 MMAs, DMMA pipe 38 % busy.
// The base rows (Z, chi, X) of the next chunk are fetched into a register while the current one is consumed.
template <int NTB>
__global__ void __launch_bounds__(RS_THREADS, 2) ragged_stats_kernel(const RaggedStatsArgs a, int gy, int gz) {
  constexpr int TILES = RS_MTB * NTB;
  __shared__ double2 s_av[2][RS_MTB][RS_THREADS];
  __shared__ double s_w[2][RS_QMAX + 1][8];    // row q stays zero: the weight row of "no pair"
  __shared__ double s_base[2][RS_BASEMAX][8];
  __shared__ uchar4 s_feat[RS_QMAX];          // feature -> rows of s_base: (Z row, chi row or 255, X row or 255)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int rows = a.bw * a.P;
  const int sid = blockIdx.y * RS_WARPS + warp;            // this warp's slice
  const bool active = sid < gy * gz;
  const int ys = sid % gy, zs = sid / gy;
  if (threadIdx.x < a.q) {
    const int f = threadIdx.x;
    const int dd = f % (1 + a.D), km = f / (1 + a.D), mm = km % (a.M + 1), k = km / (a.M + 1);
    s_feat[f] = make_uchar4((unsigned char)k, mm > 0 ? (unsigned char)(a.K + mm - 1) : 255,
                            dd > 0 ? (unsigned char)(a.K + a.M + dd - 1) : 255, 0);
  }
  if (threadIdx.x < 16) s_w[threadIdx.x >> 3][a.q][threadIdx.x & 7] = 0.0;
  const int nbase = a.K + a.M + a.D;
  // pair (fa <= fb) of this lane in each pair-tile: index -> row-major upper triangle; kept as the byte offsets of
  // the two weight rows in s_w (-1: no pair), so the hot loop forms an address with one add
  int pa[NTB], pb[NTB];
#pragma unroll
  for (int nt = 0; nt < NTB; nt++) {
    int pf = (zs * NTB + nt) * 8 + g, fa = 0;
    pa[nt] = (a.q * 8 + 2 * c) * 8; pb[nt] = pa[nt];       // the zero row: no branch in the hot loop
    if (active && pf < a.npairs) {
      int rem = pf;
      while (rem >= a.q - fa) { rem -= a.q - fa; fa++; }
      pa[nt] = (fa * 8 + 2 * c) * 8; pb[nt] = ((fa + rem) * 8 + 2 * c) * 8;
    }
  }
  const double* ap[RS_MTB];
#pragma unroll
  for (int mt = 0; mt < RS_MTB; mt++) {
    int e = (ys * RS_MTB + mt) * 8 + g;
    ap[mt] = (active && e < rows) ? a.Gl + (size_t)e * a.ld : nullptr;
  }
  double R[RS_MTB][NTB][2];
#pragma unroll
  for (int mt = 0; mt < RS_MTB; mt++)
#pragma unroll
    for (int nt = 0; nt < NTB; nt++) { R[mt][nt][0] = 0; R[mt][nt][1] = 0; }

  // staging role of this thread: element (row, slot) of the base rows
  const int srow = threadIdx.x >> 3, slot = threadIdx.x & 7;
  const double* ssrc = nullptr;
  if (srow < nbase)
    ssrc = srow < a.K ? a.Z + (size_t)srow * a.ld : (srow < a.K + a.M ? a.chi + (size_t)(srow - a.K) * a.ld : a.X + (size_t)(srow - a.K - a.M) * a.ld);
  const int n_chunks = a.ld >> 3;
  int ch = blockIdx.x;
  double nextv = (ssrc && ch < n_chunks) ? ssrc[(ch << 3) + slot] : 0.0;
  // The band rows (A operand) travel global -> shared with cp.async (LDGSTS, 16 bytes per thread, L1 bypassed) into
  // thread-private slots, one chunk AHEAD: the ~600 cycles of global-load latency hide behind the ~650 cycles of
  // tensor-pipe work of the current chunk instead of stalling every chunk, and nothing waits in registers.
  auto issue = [&](int stage, int chunk) {
    if (chunk < n_chunks) {
#pragma unroll
      for (int mt = 0; mt < RS_MTB; mt++)
        if (ap[mt]) rs_cp_async16(&s_av[stage][mt][threadIdx.x], ap[mt] + (chunk << 3) + 2 * c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0, ch);
  __syncthreads();
  for (int it = 0; ch < n_chunks; ch += gridDim.x, it++) {
    const int buf = it & 1;
    if (ssrc) s_base[buf][srow][slot] = nextv;
    const int chn = ch + gridDim.x;
    issue(buf ^ 1, chn);
    if (ssrc && chn < n_chunks) nextv = ssrc[(chn << 3) + slot];        // in flight while this chunk is consumed
    asm volatile("cp.async.wait_group 1;" ::: "memory");                // the group of the current chunk has landed
    double2 av[RS_MTB];
#pragma unroll
    for (int mt = 0; mt < RS_MTB; mt++) av[mt] = ap[mt] ? s_av[buf][mt][threadIdx.x] : make_double2(0.0, 0.0);
    __syncthreads();
    for (int f = srow; f < a.q; f += RS_THREADS / 8) {                  // feature weights, once per block
      const uchar4 t = s_feat[f];
      double w = s_base[buf][t.x][slot];
      if (t.y != 255) w *= s_base[buf][t.y][slot];
      if (t.z != 255) w *= s_base[buf][t.z][slot];
      s_w[buf][f][slot] = w;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int nt = 0; nt < NTB; nt++) {
        const char* wbase = reinterpret_cast<const char*>(&s_w[buf][0][0]);
        const double2 wa = *reinterpret_cast<const double2*>(wbase + pa[nt]);
        const double2 wb = *reinterpret_cast<const double2*>(wbase + pb[nt]);
        const double2 wv = make_double2(wa.x * wb.x, wa.y * wb.y);
#pragma unroll
        for (int mt = 0; mt < RS_MTB; mt++) {
          dmma884r(R[mt][nt][0], R[mt][nt][1], av[mt].x, wv.x);
          dmma884r(R[mt][nt][0], R[mt][nt][1], av[mt].y, wv.y);
        }
      }
    }
  }
  if (!active) return;
  // every warp owns its slice: straight from the accumulators to this block's partial row
  double* row = a.partials + (((size_t)zs * gy + ys) * gridDim.x + blockIdx.x) * (TILES * 64);
  int t = 0;
#pragma unroll
  for (int mt = 0; mt < RS_MTB; mt++)
#pragma unroll
    for (int nt = 0; nt < NTB; nt++, t++) {
      const int idx = t * 64 + g * 8 + 2 * c;
      row[idx] = R[mt][nt][0]; row[idx + 1] = R[mt][nt][1];
    }
}

// one warp per output element Hb[pair][e]
template <int NTB>
__global__ void __launch_bounds__(256) ragged_stats_final_kernel(const RaggedStatsArgs a, int gx, int gy) {
  constexpr int TILES = RS_MTB * NTB;
  const int rows = a.bw * a.P;
  const int64_t el = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (el >= (int64_t)a.npairs * rows) return;
  const int pair = (int)(el / rows), e = (int)(el % rows);
  const int bz = pair / (NTB * 8), nt = (pair % (NTB * 8)) >> 3, col = pair & 7;
  const int mtg = e >> 3, by = mtg / RS_MTB, mt = mtg % RS_MTB;
  const int idx = (mt * NTB + nt) * 64 + (e & 7) * 8 + col;
  const double* base = a.partials + (((size_t)bz * gy + by) * gx) * (TILES * 64) + idx;
  double t = 0;
  for (int b = lane; b < gx; b += 32) t += base[(size_t)b * (TILES * 64)];
  t = warp_sum(t);
  if (lane == 0) a.Hb[el] = t;
}

static void rs_shape(int P, int bw, int q, int sm_count, int& NTB, int& gx, int& gy, int& gz, int& npairs) {
  npairs = q * (q + 1) / 2;
  int ptiles = (npairs + 7) / 8;
  NTB = ptiles <= 2 ? 2 : (ptiles <= 5 ? 5 : 10);
  gz = (ptiles + NTB - 1) / NTB;
  int mtiles = (bw * P + 7) / 8;
  gy = (mtiles + RS_MTB - 1) / RS_MTB;
  const int yblocks = (gy * gz + RS_WARPS - 1) / RS_WARPS;     // blocks of eight slices
  gx = (2 * sm_count) / yblocks;         // two blocks per SM, ONE wave
  if (gx < 1) gx = 1;
}

size_t ragged_stats_partial_doubles(int P, int bw, int q, int sm_count) {
  int NTB, gx, gy, gz, np;
  rs_shape(P, bw, q, sm_count, NTB, gx, gy, gz, np);
  return (size_t)gx * gy * gz * RS_MTB * NTB * 64;
}

template <int NTB>
static int launch_rs(const RaggedStatsArgs& a, int gx, int gy, int gz, cudaStream_t s) {
  dim3 grid(gx, (gy * gz + RS_WARPS - 1) / RS_WARPS);
  ragged_stats_kernel<NTB><<<grid, RS_THREADS, 0, s>>>(a, gy, gz);
  int64_t tot = (int64_t)a.npairs * a.bw * a.P;
  ragged_stats_final_kernel<NTB><<<(unsigned)((tot + 7) / 8), 256, 0, s>>>(a, gx, gy);
  g_launch_count += 2;
  return (int)cudaGetLastError();
}

int launch_ragged_stats(const RaggedStatsArgs& a, int sm_count, cudaStream_t s) {
  if (a.q > RS_QMAX) return -6;
  int NTB, gx, gy, gz, np;
  rs_shape(a.P, a.bw, a.q, sm_count, NTB, gx, gy, gz, np);
  if (NTB == 2) return launch_rs<2>(a, gx, gy, gz, s);
  if (NTB == 5) return launch_rs<5>(a, gx, gy, gz, s);
  return launch_rs<10>(a, gx, gy, gz, s);
}

}  // namespace bf
