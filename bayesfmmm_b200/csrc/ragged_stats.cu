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

template <int NTB>
__global__ void __launch_bounds__(RS_THREADS, 2) ragged_stats_kernel(const RaggedStatsArgs a) {
  constexpr int TILES = RS_MTB * NTB;
  __shared__ double s_acc[TILES * 64];
  __shared__ double s_w[RS_WARPS][RS_QMAX][8];
  __shared__ double s_base[RS_WARPS][RS_BASEMAX][8];
  __shared__ uchar4 s_feat[RS_QMAX];          // feature -> rows of s_base: (Z row, chi row or 255, X row or 255)
  __shared__ unsigned char s_pa[NTB * 8], s_pb[NTB * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int rows = a.bw * a.P;
  const int pair0 = blockIdx.z * NTB * 8;
  // pair table of this block's slice: index -> (a, b), a <= b, row-major upper triangle
  if (threadIdx.x < NTB * 8) {
    int pf = pair0 + threadIdx.x, fa = 0;
    unsigned char pa = 255, pb = 255;
    if (pf < a.npairs) {
      int rem = pf;
      while (rem >= a.q - fa) { rem -= a.q - fa; fa++; }
      pa = (unsigned char)fa; pb = (unsigned char)(fa + rem);
    }
    s_pa[threadIdx.x] = pa; s_pb[threadIdx.x] = pb;
  }
  if (threadIdx.x < a.q) {
    const int f = threadIdx.x;
    const int dd = f % (1 + a.D), km = f / (1 + a.D), mm = km % (a.M + 1), k = km / (a.M + 1);
    s_feat[f] = make_uchar4((unsigned char)k, mm > 0 ? (unsigned char)(a.K + mm - 1) : 255,
                            dd > 0 ? (unsigned char)(a.K + a.M + dd - 1) : 255, 0);
  }
  const int nbase = a.K + a.M + a.D;
  __syncthreads();
  const double* ap[RS_MTB];
#pragma unroll
  for (int mt = 0; mt < RS_MTB; mt++) {
    int e = (blockIdx.y * RS_MTB + mt) * 8 + g;
    ap[mt] = (e < rows) ? a.Gl + (size_t)e * a.ld : nullptr;
  }
  int pa[NTB], pb[NTB];
#pragma unroll
  for (int nt = 0; nt < NTB; nt++) { pa[nt] = s_pa[nt * 8 + g]; pb[nt] = s_pb[nt * 8 + g]; }

  double R[RS_MTB][NTB][2];
#pragma unroll
  for (int mt = 0; mt < RS_MTB; mt++)
#pragma unroll
    for (int nt = 0; nt < NTB; nt++) { R[mt][nt][0] = 0; R[mt][nt][1] = 0; }

  const int n_chunks = a.ld >> 3;
  const int wstride = gridDim.x * RS_WARPS;
  const int slot = lane & 7;
  for (int ch = blockIdx.x * RS_WARPS + warp; ch < n_chunks; ch += wstride) {
    const int i8 = ch << 3;
    double2 av[RS_MTB];
#pragma unroll
    for (int mt = 0; mt < RS_MTB; mt++)
      av[mt] = ap[mt] ? __ldcs(reinterpret_cast<const double2*>(ap[mt] + i8 + 2 * c)) : make_double2(0.0, 0.0);
    // feature weights of the 8 functions of this chunk: w[f][slot].  The K + M + D base rows are staged
    // once (two loads per lane), the products are formed from shared memory with the feature -> (k, m, d)
    // table built at block start (no integer division in the loop).
    for (int r = lane >> 3; r < nbase; r += 4) {
      const double* src = r < a.K ? a.Z + (size_t)r * a.ld : (r < a.K + a.M ? a.chi + (size_t)(r - a.K) * a.ld : a.X + (size_t)(r - a.K - a.M) * a.ld);
      s_base[warp][r][slot] = src[i8 + slot];
    }
    __syncwarp();
    for (int f = lane >> 3; f < a.q; f += 4) {
      const uchar4 t = s_feat[f];
      double w = s_base[warp][t.x][slot];
      if (t.y != 255) w *= s_base[warp][t.y][slot];
      if (t.z != 255) w *= s_base[warp][t.z][slot];
      s_w[warp][f][slot] = w;
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NTB; nt++) {
      double2 wv = make_double2(0.0, 0.0);
      if (pa[nt] != 255) {
        const double2 wa = *reinterpret_cast<const double2*>(&s_w[warp][pa[nt]][2 * c]);
        const double2 wb = *reinterpret_cast<const double2*>(&s_w[warp][pb[nt]][2 * c]);
        wv.x = wa.x * wb.x; wv.y = wa.y * wb.y;
      }
#pragma unroll
      for (int mt = 0; mt < RS_MTB; mt++) {
        dmma884r(R[mt][nt][0], R[mt][nt][1], av[mt].x, wv.x);
        dmma884r(R[mt][nt][0], R[mt][nt][1], av[mt].y, wv.y);
      }
    }
    __syncwarp();
  }
  for (int w = 0; w < RS_WARPS; w++) {
    if (warp == w) {
      int t = 0;
#pragma unroll
      for (int mt = 0; mt < RS_MTB; mt++)
#pragma unroll
        for (int nt = 0; nt < NTB; nt++, t++) {
          int idx = t * 64 + g * 8 + 2 * c;
          if (w == 0) { s_acc[idx] = R[mt][nt][0]; s_acc[idx + 1] = R[mt][nt][1]; }
          else { s_acc[idx] += R[mt][nt][0]; s_acc[idx + 1] += R[mt][nt][1]; }
        }
    }
    __syncthreads();
  }
  double* row = a.partials + (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (TILES * 64);
  for (int idx = threadIdx.x; idx < TILES * 64; idx += RS_THREADS) row[idx] = s_acc[idx];
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
  gx = (2 * sm_count) / (gy * gz);       // two blocks per SM, ONE wave (rounding up left a 6 % second wave: 2x the time)
  if (gx < 1) gx = 1;
}

size_t ragged_stats_partial_doubles(int P, int bw, int q, int sm_count) {
  int NTB, gx, gy, gz, np;
  rs_shape(P, bw, q, sm_count, NTB, gx, gy, gz, np);
  return (size_t)gx * gy * gz * RS_MTB * NTB * 64;
}

template <int NTB>
static int launch_rs(const RaggedStatsArgs& a, int gx, int gy, int gz, cudaStream_t s) {
  dim3 grid(gx, gy, gz);
  ragged_stats_kernel<NTB><<<grid, RS_THREADS, 0, s>>>(a);
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
