// ragged_prep.cu -- create-time projection for RAGGED per-function grids (BASELINE config 4).
//
// With its own grid every function has its own Gram G_i = B_i'B_i, so there is no common whitening.
// What the per-iteration kernels need is still data-only:
//     ||y_i - B_i theta||^2 = rss_i + (c_i - theta)' G_i (c_i - theta)
// for ANY least-squares solution c_i (G_i c_i = B_i'y_i), with rss_i = ||y_i - B_i c_i||^2.
// For a B-spline basis G_i is banded (bandwidth degree + 1), so the cache per function is
//     c_i (P) | h_i = B_i'y_i (P) | the lower band of G_i (bw x P) | rss_i
// = (2 + bw) P + 1 doubles (101 for cubic splines, P = 20) instead of y_i and t_i (2 n_i ~ 400).
// One thread per function; scratch in local memory (one-time kernel, ~n_i (degree+1)^2 flops).
#include "common.cuh"

namespace bf {

constexpr int R_PMAX = 64;
constexpr int R_BWMAX = 8;

__device__ inline int bspline_eval_point(double x, const double* __restrict__ kn, int nk, int degree, double* h) {
  // returns the first column of the degree+1 non-zero basis values written to h, or -1 outside the knots
  if (x < kn[0] || x > kn[nk - 1]) return -1;
  int ell = degree;
  while (ell < nk - degree - 2 && x >= kn[ell + 1]) ell++;
  double hh[R_BWMAX];
  h[0] = 1.0;
  for (int j = 1; j <= degree; j++) {
    for (int q = 0; q < j; q++) hh[q] = h[q];
    h[0] = 0.0;
    for (int q = 1; q <= j; q++) {
      double xb = kn[ell + q], xa = kn[ell + q - j];
      if (xb == xa) { h[q] = 0.0; continue; }
      double w = hh[q - 1] / (xb - xa);
      h[q - 1] = __dadd_rn(h[q - 1], __dmul_rn(w, xb - x));
      h[q] = __dmul_rn(w, x - xa);
    }
  }
  return ell - degree;
}

// band width of user-supplied basis rows: max over rows of (last non-zero - first non-zero + 1)
__global__ void band_width_kernel(const double* __restrict__ B, int64_t rows, int P, int* __restrict__ bw) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const double* b = B + r * P;
  int first = -1, last = -1;
  for (int p = 0; p < P; p++)
    if (b[p] != 0.0) { if (first < 0) first = p; last = p; }
  if (first >= 0) atomicMax(bw, last - first + 1);
}

__global__ void __launch_bounds__(64) ragged_prep_kernel(const RaggedPrepArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int P = a.P, bw = a.bw;
  double Gl[R_PMAX * R_BWMAX];   // Gl[j*P + p] = G[p-j][p]
  double h[R_PMAX], c[R_PMAX], w[R_BWMAX];
  for (int e = 0; e < P * bw; e++) Gl[e] = 0.0;
  for (int p = 0; p < P; p++) h[p] = 0.0;
  const int64_t lo = a.off[i], hi = a.off[i + 1];
  // ---- pass 1: banded Gram and B'y
  for (int64_t l = lo; l < hi; l++) {
    int first;
    if (a.Brows) {
      const double* b = a.Brows + l * P;
      first = -1;
      int last = -1;
      for (int p = 0; p < P; p++)
        if (b[p] != 0.0) { if (first < 0) first = p; last = p; }
      if (first < 0) continue;
      for (int q = 0; q < bw; q++) w[q] = (first + q <= last) ? b[first + q] : 0.0;
    } else {
      first = bspline_eval_point(a.t[l], a.knots, a.n_knots, a.degree, w);
      if (first < 0) continue;
    }
    const double yl = a.y[l];
    for (int q = 0; q < bw; q++) {
      const int p = first + q;
      if (p >= P) break;
      h[p] = fma(w[q], yl, h[p]);
      for (int j = 0; j <= q; j++) Gl[j * P + p] = fma(w[q - j], w[q], Gl[j * P + p]);   // G[p-j][p]
    }
  }
  // ---- banded Cholesky G = L L' (L lower band, Lb[j*P + p] = L[p][p-j]); tiny pivots are skipped so
  // that a rank-deficient G_i (no observation in some knot span) still yields a least-squares solution
  double Lb[R_PMAX * R_BWMAX];
  double dmax = 0.0;
  for (int p = 0; p < P; p++) dmax = fmax(dmax, Gl[p]);
  const double tol = dmax * 1e-13;
  for (int p = 0; p < P; p++) {
    for (int j = bw - 1; j >= 0; j--) {          // column k = p - j of row p
      const int k = p - j;
      if (k < 0) { Lb[j * P + p] = 0.0; continue; }
      double s = Gl[j * P + p];                  // G[k][p]
      for (int m = 1; m < bw - j; m++) {         // sum over columns q = k - m shared by rows p and k
        const int q = k - m;
        if (q < 0) break;
        s -= Lb[(j + m) * P + p] * Lb[m * P + k];
      }
      if (j == 0) Lb[p] = (s > tol) ? sqrt(s) : 0.0;
      else Lb[j * P + p] = (Lb[k] > 0.0) ? s / Lb[k] : 0.0;
    }
  }
  // forward L z = h, backward L' c = z
  for (int p = 0; p < P; p++) {
    double s = h[p];
    for (int j = 1; j < bw && p - j >= 0; j++) s -= Lb[j * P + p] * c[p - j];
    c[p] = (Lb[p] > 0.0) ? s / Lb[p] : 0.0;
  }
  for (int p = P - 1; p >= 0; p--) {
    double s = c[p];
    for (int j = 1; j < bw && p + j < P; j++) s -= Lb[j * P + (p + j)] * c[p + j];
    c[p] = (Lb[p] > 0.0) ? s / Lb[p] : 0.0;
  }
  // ---- pass 2: orthogonal residual, evaluated directly
  double rss = 0.0;
  for (int64_t l = lo; l < hi; l++) {
    double fit = 0.0;
    if (a.Brows) {
      const double* b = a.Brows + l * P;
      for (int p = 0; p < P; p++) fit = fma(b[p], c[p], fit);
    } else {
      int first = bspline_eval_point(a.t[l], a.knots, a.n_knots, a.degree, w);
      if (first >= 0)
        for (int q = 0; q < bw && first + q < P; q++) fit = fma(w[q], c[first + q], fit);
    }
    const double r = a.y[l] - fit;
    rss = fma(r, r, rss);
  }
  const size_t col = (size_t)a.i_begin + i;
  for (int p = 0; p < P; p++) {
    a.C[(size_t)p * a.ld + col] = c[p];
    a.H[(size_t)p * a.ld + col] = h[p];
  }
  for (int e = 0; e < P * bw; e++) a.Gl[(size_t)e * a.ld + col] = Gl[e];
  a.rss[col] = rss;
}

int launch_band_width(const double* B, int64_t rows, int P, int* bw_dev, cudaStream_t s) {
  band_width_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(B, rows, P, bw_dev);
  g_launch_count++;
  return (int)cudaGetLastError();
}

int launch_ragged_prep(const RaggedPrepArgs& a, cudaStream_t s) {
  if (a.P > R_PMAX || a.bw > R_BWMAX) return -5;
  ragged_prep_kernel<<<(a.n + 63) / 64, 64, 0, s>>>(a);
  g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace bf
