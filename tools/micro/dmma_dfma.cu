// Do FP64 DMMA (mma.m8n8k4.f64) and DFMA share a pipe on B200?  Times N iterations of (a) 8 independent
// DMMA chains, (b) 16 independent DFMA chains, (c) both interleaved, per warp, with 8 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, int iters, double a, double b) {
  double m[8][2], f[16];
  for (int i = 0; i < 8; i++) { m[i][0] = threadIdx.x; m[i][1] = i; }
  for (int i = 0; i < 16; i++) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma(m[i][0], m[i][1], a, b);
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += m[i][0] + m[i][1];
  for (int i = 0; i < 16; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(double* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(d, 10, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1); if (cudaGetLastError() != cudaSuccess) printf("launch error\n");
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 1024 * 8);
  const int iters = 20000;
  float a = run<1>(d, iters), b = run<2>(d, iters), c = run<3>(d, iters);
  // per SM sub-partition: 8 warps x iters x (8 DMMA | 16 DFMA)
  double cyc = 1.965e6;  // cycles per ms at 1965 MHz
  printf("dmma only  %.3f ms  -> %.2f cycles per DMMA per sub-partition\n", a, a * cyc / (4.0 * iters * 8));
  printf("dfma only  %.3f ms  -> %.2f cycles per DFMA per sub-partition\n", b, b * cyc / (4.0 * iters * 16));
  printf("both       %.3f ms  (sum %.3f, max %.3f): %s\n", c, a + b, a > b ? a : b, c > 0.9 * (a + b) ? "SHARED pipe" : "separate pipes (overlap)");
  return 0;
}
