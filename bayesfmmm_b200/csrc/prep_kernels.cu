// prep_kernels.cu -- create-time kernels: device B-spline evaluation and the projection of the raw
// observations onto the whitened basis (the cache the per-iteration passes stream).
#include "common.cuh"

namespace bf {

// ------------------------------------------------------------------ B-spline design matrix
// Replaces splines2::BSpline(t, internal_knots, degree, boundary_knots).basis(true) at
// reference BFMMM.h:1188-1196 / UserFunctions.cpp:820-831: clamped knots, intercept column kept,
// right boundary closed (last basis function = 1 there).  de Boor's recurrence; one thread per point.
constexpr int MAX_DEGREE = 7;
__global__ void bspline_kernel(const double* __restrict__ t, int64_t n, const double* __restrict__ kn,
                               int nk, int degree, int P, double* __restrict__ B) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double x = t[r];
  double* row = B + r * P;
  for (int p = 0; p < P; p++) row[p] = 0.0;
  if (x < kn[0] || x > kn[nk - 1]) return;
  int ell = degree;
  while (ell < nk - degree - 2 && x >= kn[ell + 1]) ell++;
  double h[MAX_DEGREE + 1], hh[MAX_DEGREE + 1];
  h[0] = 1.0;
  for (int j = 1; j <= degree; j++) {
    for (int q = 0; q < j; q++) hh[q] = h[q];
    h[0] = 0.0;
    for (int q = 1; q <= j; q++) {
      double xb = kn[ell + q], xa = kn[ell + q - j];
      if (xb == xa) { h[q] = 0.0; continue; }
      double w = hh[q - 1] / (xb - xa);
      h[q - 1] = __dadd_rn(h[q - 1], __dmul_rn(w, xb - x));   // no FMA contraction: bit-equal to the CPU recurrence
      h[q] = __dmul_rn(w, x - xa);
    }
  }
  for (int q = 0; q <= degree; q++) row[ell - degree + q] = h[q];
}

int launch_bspline(const double* t, int64_t n, const double* knots, int n_knots, int degree, int P,
                   double* B, cudaStream_t s) {
  if (degree > MAX_DEGREE) return -4;
  bspline_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(t, n, knots, n_knots, degree, P, B);
  g_launch_count++;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ statistics read-back without a copy engine
// The sampler reads a few KB of reduced statistics two or three times per sweep.  As cudaMemcpyAsync
// those reads queue on the device-to-host copy engine behind an overlapped 48 MB state transfer
// (bfmmm_get_state_begin) and stall the sweep; stored by SM threads into mapped page-locked memory
// they do not.  dst is the device alias of the host buffer.
__global__ void __launch_bounds__(256) copy_to_host_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t len) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __ldcg(src + i);
  __threadfence_system();
}
int launch_copy_to_host(const double* src, double* dst_mapped, int64_t len, cudaStream_t s) {
  if (len <= 0) return 0;
  int64_t blocks = (len + 255) / 256;
  if (blocks > 64) blocks = 64;
  copy_to_host_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dst_mapped, len);
  g_launch_count++;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ sigma^2 on the device
// updateSigma's draw (UpdateSigma.h:47-53; tempered :98-107): sigma^2 = 1 / ((1 / b1) Gamma(a)),
// b1 = scale * SSR + beta_0, from the Philox stream (key, 0xB200, iteration, purpose) -- the stream and the
// sampler (RngStream::gamma) the host-side draw uses.  One thread, right behind the SSR pass (and its
// all-reduce): the chi kernel that follows reads sigma^2 from device memory, so the sweep does not wait
// for a host round trip between the two passes.  (SSR, sigma^2) reach the host through mapped
// page-locked memory; the sequence number is stored last.
__global__ void sigma_draw_kernel(const double* __restrict__ ssr_dev, double a, double scale_ssr, double beta0, uint64_t key,
                                  uint64_t iteration, uint32_t purpose, double* __restrict__ sigma_dev,
                                  volatile double* host, double seq) {
  RngStream rs(key, 0xB200ull, iteration, purpose);
  const double ssr = *ssr_dev;
  const double b1 = scale_ssr * ssr + beta0;
  const double r = (1 / b1) * rs.gamma(a);
  const double sig = 1 / r;
  *sigma_dev = sig;
  host[0] = ssr; host[1] = sig;
  __threadfence_system();
  host[2] = seq;
}
int launch_sigma_draw(const double* ssr_dev, double a, double scale_ssr, double beta0, uint64_t key, uint64_t iteration,
                      uint32_t purpose, double* sigma_dev, double* host_mapped, double seq, cudaStream_t s) {
  sigma_draw_kernel<<<1, 1, 0, s>>>(ssr_dev, a, scale_ssr, beta0, key, iteration, purpose, sigma_dev, host_mapped, seq);
  g_launch_count++;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ projection
// c~_i = Q' y_i with Q = B L^{-T} (orthonormal columns), rss_i = ||y_i - Q c~_i||^2 evaluated
// directly (not as ||y||^2 - ||c~||^2, which cancels catastrophically when sigma^2 << signal).
// One warp per function: lanes stride over the grid points (coalesced reads of y_i), P is tiled
// in chunks of PCH coefficients held in registers.
constexpr int PCH = 16;
constexpr int PJ_WARPS = 8;
__global__ void __launch_bounds__(PJ_WARPS * 32) project_kernel(const ProjectArgs a) {
  extern __shared__ double sm[];            // per warp: P coefficients
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cw = sm + (size_t)warp * a.P;
  const int64_t fn = (int64_t)blockIdx.x * PJ_WARPS + warp;
  if (fn >= a.n) return;
  const double* y = a.Y + fn * a.T;
  for (int p0 = 0; p0 < a.P; p0 += PCH) {
    double acc[PCH];
#pragma unroll
    for (int j = 0; j < PCH; j++) acc[j] = 0;
    for (int64_t t = lane; t < a.T; t += 32) {
      double yv = y[t];
      const double* q = a.Q + t * a.P + p0;
#pragma unroll
      for (int j = 0; j < PCH; j++)
        if (p0 + j < a.P) acc[j] = fma(yv, q[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < PCH; j++) {
      double v = warp_sum(acc[j]);
      if (lane == 0 && p0 + j < a.P) cw[p0 + j] = v;
    }
  }
  __syncwarp();
  double rs = 0;
  for (int64_t t = lane; t < a.T; t += 32) {
    const double* q = a.Q + t * a.P;
    double fit = 0;
    for (int p = 0; p < a.P; p++) fit = fma(q[p], cw[p], fit);
    double r = y[t] - fit;
    rs = fma(r, r, rs);
  }
  rs = warp_sum(rs);
  const int64_t col = a.i_begin + fn;
  for (int p = lane; p < a.P; p += 32) a.Ct[(size_t)p * a.ld + col] = cw[p];
  if (lane == 0) a.rss[col] = rs;
}

int launch_project(const ProjectArgs& a, cudaStream_t s) {
  size_t smem = (size_t)PJ_WARPS * a.P * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned blocks = (unsigned)((a.n + PJ_WARPS - 1) / PJ_WARPS);
  project_kernel<<<blocks, PJ_WARPS * 32, smem, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace bf
