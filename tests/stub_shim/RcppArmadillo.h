// Test infrastructure: the few Armadillo / Rcpp / Rmath names the reference-side stub of INTEGRATION.md section 3 uses,
// with Armadillo's memory layout (column-major, cubes contiguous slice after slice), so that the stub can be compiled
// and linked against include/bfmmm.h + libbfmmm_b200.so where R, Rcpp and Armadillo are absent
// (tests/test_integration_stub.py).  Not a port of Armadillo and not used by the product.
#pragma once
#include <cstddef>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace arma {
typedef unsigned long long uword;
class vec;
struct subvec_ref {
  double* p; uword n;
  subvec_ref& operator=(const vec& v);
};
class vec {
 public:
  std::vector<double> mem; uword n_elem = 0, n_rows = 0;
  vec() {}
  explicit vec(uword n) : mem(n, 0.0), n_elem(n), n_rows(n) {}
  double* memptr() { return mem.data(); }
  const double* memptr() const { return mem.data(); }
  double& operator()(uword i) { return mem[i]; }
  const double& operator()(uword i) const { return mem[i]; }
  subvec_ref subvec(uword a, uword b) { return subvec_ref{mem.data() + a, b - a + 1}; }
};
inline subvec_ref& subvec_ref::operator=(const vec& v) { for (uword i = 0; i < n; i++) p[i] = v.mem[i]; return *this; }
class mat {
 public:
  std::vector<double> own; double* ptr = nullptr; uword n_rows = 0, n_cols = 0, n_elem = 0;
  mat() {}
  mat(uword r, uword c) : own(r * c, 0.0), n_rows(r), n_cols(c), n_elem(r * c) { ptr = own.data(); }
  mat(double* p, uword r, uword c) : ptr(p), n_rows(r), n_cols(c), n_elem(r * c) {}        // a view (cube slice)
  double* memptr() { return ptr; }
  const double* memptr() const { return ptr; }
  double& operator()(uword r, uword c) { return ptr[c * n_rows + r]; }
  const double& operator()(uword r, uword c) const { return ptr[c * n_rows + r]; }
  mat& operator=(const mat& o) {                                                            // element copy, also into a view
    if (!ptr) { own.assign(o.n_elem, 0.0); ptr = own.data(); n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; }
    for (uword i = 0; i < n_elem; i++) ptr[i] = o.ptr[i];
    return *this;
  }
  mat(const mat& o) : own(o.ptr, o.ptr + o.n_elem), n_rows(o.n_rows), n_cols(o.n_cols), n_elem(o.n_elem) { ptr = own.data(); }
};
class cube {
 public:
  std::vector<double> mem; uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
  cube() {}
  cube(uword r, uword c, uword s) : mem(r * c * s, 0.0), n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s) {}
  double* memptr() { return mem.data(); }
  const double* memptr() const { return mem.data(); }
  mat slice(uword s) { return mat(mem.data() + s * n_rows * n_cols, n_rows, n_cols); }
  double& operator()(uword r, uword c, uword s) { return mem[(s * n_cols + c) * n_rows + r]; }
};
template <class T>
class field {
 public:
  std::vector<T> items; uword n_rows = 0, n_cols = 0;
  field() {}
  field(uword r, uword c) : items(r * c), n_rows(r), n_cols(c) {}
  T& operator()(uword r, uword c) { return items[c * n_rows + r]; }
  const T& operator()(uword r, uword c) const { return items[c * n_rows + r]; }
};
}  // namespace arma

namespace Rcpp {
struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] inline void stop(const char* msg) { throw exception(msg ? msg : "error"); }
}  // namespace Rcpp

namespace R {
inline double rnorm(double mu, double sd) {
  static std::mt19937_64 gen(12345);
  return std::normal_distribution<double>(mu, sd)(gen);
}
}  // namespace R
