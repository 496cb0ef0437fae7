// Dependent-issue latency of FP64 / FP32 / integer instructions and shared-memory loads on one warp (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double x, long long* out, double* sink) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x + threadIdx.x;
  __syncthreads();
  double a = x, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < 1024; i++) a = fma(a, b, c);
  long long t1 = clock64();
  float f = (float)x, g = 1.0000001f;
#pragma unroll 64
  for (int i = 0; i < 1024; i++) f = fmaf(f, g, 1e-9f);
  long long t2 = clock64();
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
#pragma unroll 16
  for (int i = 0; i < 256; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
  long long t3 = clock64();
  int idx = threadIdx.x & 63;
#pragma unroll 16
  for (int i = 0; i < 256; i++) idx = (int)sm[idx & 63] & 63;
  long long t4 = clock64();
  double r = x + 3.0;
#pragma unroll 16
  for (int i = 0; i < 256; i++) r = rsqrt(r) + 2.0;
  long long t5 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; }
  sink[threadIdx.x] = a + f + a0 + a1 + a2 + a3 + idx + r;
}
int main() {
  long long* d; double* s; cudaMalloc(&d, 64); cudaMalloc(&s, 8 * 32);
  lat<<<1, 32>>>(1.5, d, s); cudaDeviceSynchronize();
  lat<<<1, 32>>>(1.5, d, s); cudaDeviceSynchronize();
  long long h[5]; cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
  printf("dependent DFMA: %.1f cycles each; dependent FFMA: %.1f; 4 independent DFMA chains: %.1f cycles per DFMA; dependent LDS.64+cvt: %.1f; dependent rsqrt(double)+add: %.1f\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 256.0, h[4] / 256.0);
  return 0;
}
