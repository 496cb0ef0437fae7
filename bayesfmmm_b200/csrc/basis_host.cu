// basis_host.cu -- host-side basis / penalty construction (include/bfmmm_basis.h).
#include <cmath>
#include <string>
#include <vector>

#include "../../include/bfmmm_basis.h"
#include "common.cuh"

namespace {
int bfail(const std::string& m) { return bf::set_error(m.c_str()); }

// de Boor's recurrence for the degree+1 non-zero basis functions at x; returns their first column
int basis_funs(double x, const std::vector<double>& kn, int degree, double* h, double* hh) {
  const int nk = (int)kn.size();
  if (x < kn[0] || x > kn[nk - 1]) return -1;
  int ell = degree;
  while (ell < nk - degree - 2 && x >= kn[ell + 1]) ell++;
  h[0] = 1.0;
  for (int j = 1; j <= degree; j++) {
    for (int q = 0; q < j; q++) hh[q] = h[q];
    h[0] = 0.0;
    for (int q = 1; q <= j; q++) {
      const double xb = kn[ell + q], xa = kn[ell + q - j];
      if (xb == xa) { h[q] = 0.0; continue; }
      const double w = hh[q - 1] / (xb - xa);
      h[q - 1] += w * (xb - x);
      h[q] = w * (x - xa);
    }
  }
  return ell - degree;
}

std::vector<double> clamped(const double* ik, int n_ik, int degree, double lo, double hi) {
  std::vector<double> kn(n_ik + 2 * (degree + 1));
  for (int i = 0; i <= degree; i++) { kn[i] = lo; kn[kn.size() - 1 - i] = hi; }
  for (int i = 0; i < n_ik; i++) kn[degree + 1 + i] = ik[i];
  return kn;
}
}  // namespace

extern "C" {

int bfmmm_bspline_basis(const double* t, int64_t n, const double* ik, int n_ik, int degree, double lo, double hi,
                        double* B) {
  if (!t || !B || degree < 0 || n_ik < 0 || (n_ik > 0 && !ik)) return bfail("bfmmm_bspline_basis: bad argument");
  const int P = n_ik + degree + 1;
  const std::vector<double> kn = clamped(ik, n_ik, degree, lo, hi);
  std::vector<double> h(degree + 1), hh(degree + 1);
  for (int64_t r = 0; r < n; r++) {
    double* row = B + (size_t)r * P;
    for (int p = 0; p < P; p++) row[p] = 0.0;
    const int first = basis_funs(t[r], kn, degree, h.data(), hh.data());
    if (first < 0) continue;
    for (int q = 0; q <= degree; q++) row[first + q] = h[q];
  }
  return 0;
}

int bfmmm_tensor_P(int dim, const int32_t* degree, const int32_t* n_ik) {
  int P = 1;
  for (int l = 0; l < dim; l++) P *= n_ik[l] + degree[l] + 1;
  return P;
}

// column c <-> multi-index with dimension 0 slowest (BSplines.h:29-31,55-58); B starts at ones and is
// multiplied dimension by dimension (:35,49-51)
int bfmmm_tensor_bspline(const double* t, int64_t n, int dim, const int32_t* degree, const double* boundary,
                         const double* iknots, const int32_t* n_ik, double* B) {
  if (!t || !B || dim < 1) return bfail("bfmmm_tensor_bspline: bad argument");
  std::vector<int> Pd(dim);
  int P = 1;
  for (int l = 0; l < dim; l++) { Pd[l] = n_ik[l] + degree[l] + 1; P *= Pd[l]; }
  std::vector<std::vector<double>> Bd(dim);
  const double* ik = iknots;
  for (int l = 0; l < dim; l++) {
    Bd[l].resize((size_t)n * Pd[l]);
    if (bfmmm_bspline_basis(t + (size_t)l * n, n, ik, n_ik[l], degree[l], boundary[2 * l], boundary[2 * l + 1], Bd[l].data()))
      return 1;
    ik += n_ik[l];
  }
  std::vector<int> idx(dim);
  for (int c = 0; c < P; c++) {
    int rem = c;
    for (int l = dim - 1; l >= 0; l--) { idx[l] = rem % Pd[l]; rem /= Pd[l]; }
    for (int64_t r = 0; r < n; r++) {
      double v = 1.0;
      for (int l = 0; l < dim; l++) v = v * Bd[l][(size_t)r * Pd[l] + idx[l]];
      B[(size_t)r * P + c] = v;
    }
  }
  return 0;
}

// P = C'C with one first-difference row per pair of neighbouring multi-indices (BSplines.h:96-118)
int bfmmm_get_P(int dim, const int32_t* degree, const int32_t* n_ik, double* Pm) {
  if (!Pm || dim < 1) return bfail("bfmmm_get_P: bad argument");
  std::vector<int> Pd(dim);
  int P = 1;
  for (int l = 0; l < dim; l++) { Pd[l] = n_ik[l] + degree[l] + 1; P *= Pd[l]; }
  std::vector<int> index((size_t)P * dim);
  for (int c = 0; c < P; c++) {
    int rem = c;
    for (int l = dim - 1; l >= 0; l--) { index[(size_t)c * dim + l] = rem % Pd[l]; rem /= Pd[l]; }
  }
  for (size_t e = 0; e < (size_t)P * P; e++) Pm[e] = 0.0;
  for (int i = 0; i < P; i++)
    for (int j = i; j < P; j++) {
      int diff = 0, adiff = 0;
      for (int l = 0; l < dim; l++) {
        const int dl = index[(size_t)j * dim + l] - index[(size_t)i * dim + l];
        diff += dl; adiff += dl < 0 ? -dl : dl;
      }
      if (diff == 1 && adiff == 1) {
        Pm[(size_t)i * P + i] += 1; Pm[(size_t)j * P + j] += 1;
        Pm[(size_t)j * P + i] -= 1; Pm[(size_t)i * P + j] -= 1;
      }
    }
  return 0;
}

int bfmmm_pmat_rw1(int P, double* Pm) {
  if (!Pm || P < 1) return bfail("bfmmm_pmat_rw1: bad argument");
  for (size_t e = 0; e < (size_t)P * P; e++) Pm[e] = 0.0;
  for (int j = 0; j < P; j++) {
    Pm[0] = 1;
    if (j > 0) { Pm[(size_t)j * P + j] = 2; Pm[(size_t)j * P + j - 1] = -1; Pm[(size_t)(j - 1) * P + j] = -1; }
    Pm[(size_t)(P - 1) * P + P - 1] = 1;
  }
  return 0;
}

}  // extern "C"
