// fastmath_debug.cu -- exposes the device routines of fastmath.cuh to the accuracy tests
// (include/bfmmm_debug.h: bfmmm_debug_fastmath); not used by the sampler.
#include "../../include/bfmmm_debug.h"
#include "fastmath.cuh"

namespace bf {
__global__ void fastmath_kernel(int which, const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  build_log_table();
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    double r = 0;
    switch (which) {
      case 0: r = fast_log(v); break;
      case 1: r = fast_rcp(v); break;
      case 2: r = fast_sqrt(v); break;
      case 3: r = lgamma_pos(v, fast_log(v)); break;
      case 4: { double cs, sn; sincos_u32((uint32_t)v, cs, sn); r = cs; break; }
      case 5: { double cs, sn; sincos_u32((uint32_t)v, cs, sn); r = sn; break; }
      case 6: r = log1p_series(v); break;
      case 7: { const uint64_t b = (uint64_t)v; r = u52((uint32_t)(b >> 32), (uint32_t)b); break; }   // v < 2^53: hi word has 21 bits
      case 8: { double n0, n1; fast_box_muller(0x9e3779b9u * (uint32_t)v, 0x85ebca6bu * (uint32_t)v + 1u, 0xc2b2ae35u * (uint32_t)v, n0, n1); r = n0; break; }
      case 9: { double n0, n1; fast_box_muller(0x9e3779b9u * (uint32_t)v, 0x85ebca6bu * (uint32_t)v + 1u, 0xc2b2ae35u * (uint32_t)v, n0, n1); r = n1; break; }
      default: r = v;
    }
    y[i] = r;
  }
}
}  // namespace bf

extern "C" int bfmmm_debug_fastmath(int which, const double* x, double* y, int64_t n) {
  if (n <= 0) return 0;
  double *dx = nullptr, *dy = nullptr;
  if (cudaMalloc(&dx, n * 8) != cudaSuccess || cudaMalloc(&dy, n * 8) != cudaSuccess) { cudaFree(dx); return bf::set_error("bfmmm_debug_fastmath: cudaMalloc failed"); }
  cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice);
  bf::fastmath_kernel<<<148, 256>>>(which, dx, dy, n);
  bf::g_launch_count++;
  cudaError_t e = cudaMemcpy(y, dy, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy);
  return e == cudaSuccess ? 0 : bf::set_error(cudaGetErrorString(e));
}
