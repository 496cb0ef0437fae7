// quantile_kernels.cu -- quantiles over stored draws on the device.
//
// The reference's credible-interval functions (ZCI src/PostProcessing.cpp:3505-3592, SigmaCI :3435-3480, FMeanCI
// :99-480, FCovCI :1781-2300) end in  arma::quantile(draws, {alpha/2, 1/2, 1 - alpha/2})  per element: n K elements for
// Z, T (or T^2) for the mean / covariance functions.  One thread block per element: the S draws are gathered into
// shared memory (consecutive blocks read consecutive addresses of a draw), bitonic-sorted there, and the requested
// quantiles are interpolated with Armadillo's definition (the MATLAB / Octave one, Hyndman-Fan type 5):
//   h = S p + 1/2;  q = x_(floor h) + (h - floor h) (x_(floor h + 1) - x_(floor h)),  clamped to [x_(1), x_(S)].
// Armadillo is not vendored in the reference: the definition is taken from its documentation (parity unpinned).
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/bfmmm_post.h"
#include "common.cuh"

namespace bf {
constexpr int QT_THREADS = 256;

// x: [S][R] draw-major (element r of draw s at x[s * R + r]); q: [np][R]
__global__ void __launch_bounds__(QT_THREADS) quantile_kernel(const double* __restrict__ x, int64_t S, int64_t R, int S2,
                                                              const double* __restrict__ probs, int np, double* __restrict__ q) {
  extern __shared__ double v[];                       // S2 = S rounded up to a power of two
  const int64_t r = blockIdx.x;
  for (int s = threadIdx.x; s < S2; s += QT_THREADS) v[s] = s < S ? x[(size_t)s * R + r] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= S2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < S2; i += QT_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const double a = v[i], b = v[l];
          if ((a > b) == up) { v[i] = b; v[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int p = threadIdx.x; p < np; p += QT_THREADS) {
    const double h = (double)S * probs[p] + 0.5;
    double out;
    if (h <= 1.0) out = v[0];
    else if (h >= (double)S) out = v[S - 1];
    else {
      const int64_t lo = (int64_t)floor(h);
      const double w = h - (double)lo;
      out = v[lo - 1] + w * (v[lo] - v[lo - 1]);
    }
    q[(size_t)p * R + r] = out;
  }
}
}  // namespace bf

extern "C" int bfmmm_quantiles(const double* draws, int64_t S, int64_t R, const double* probs, int np, double* out, int device) {
  if (!draws || !probs || !out || S < 1 || R < 1 || np < 1) return bf::set_error("bfmmm_quantiles: bad argument");
  int S2 = 1;
  while (S2 < S) S2 <<= 1;
  if ((size_t)S2 * 8 > 200 * 1024) return bf::set_error("bfmmm_quantiles: at most 25600 draws per element");
  if (cudaSetDevice(device) != cudaSuccess) return bf::set_error("bfmmm_quantiles: no CUDA device (no CPU fallback)");
  const size_t smem = (size_t)S2 * 8;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(bf::quantile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return bf::set_error("bfmmm_quantiles: shared memory");
  // elements in column chunks of at most 2 GB of draws
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(R, (int64_t)(1ull << 28) / S));
  double *dx = nullptr, *dq = nullptr, *dp = nullptr;
  auto fail = [&](const char* m) { cudaFree(dx); cudaFree(dq); cudaFree(dp); return bf::set_error(m); };
  if (cudaMalloc(&dx, (size_t)S * chunk * 8) != cudaSuccess || cudaMalloc(&dq, (size_t)np * chunk * 8) != cudaSuccess ||
      cudaMalloc(&dp, (size_t)np * 8) != cudaSuccess)
    return fail("bfmmm_quantiles: cudaMalloc failed");
  cudaMemcpy(dp, probs, (size_t)np * 8, cudaMemcpyHostToDevice);
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    const int64_t m = std::min<int64_t>(chunk, R - r0);
    if (cudaMemcpy2D(dx, (size_t)m * 8, draws + r0, (size_t)R * 8, (size_t)m * 8, (size_t)S, cudaMemcpyHostToDevice) != cudaSuccess)
      return fail("bfmmm_quantiles: upload failed");
    bf::quantile_kernel<<<(unsigned)m, bf::QT_THREADS, smem>>>(dx, S, m, S2, dp, np, dq);
    bf::g_launch_count++;
    if (cudaMemcpy2D(out + r0, (size_t)R * 8, dq, (size_t)m * 8, (size_t)m * 8, (size_t)np, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail((std::string("bfmmm_quantiles: ") + cudaGetErrorString(cudaGetLastError())).c_str());
  }
  cudaFree(dx); cudaFree(dq); cudaFree(dp);
  return 0;
}
