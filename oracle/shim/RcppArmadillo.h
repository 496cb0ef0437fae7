// Minimal stand-in for <RcppArmadillo.h>, written from Armadillo's and R's public
// documentation, so that the reference's own update headers
// (/root/reference/inst/include/BayesFMMM/Update*.h, CalculateLikelihood.h, ...) can be
// compiled here WITHOUT R, Rcpp or Armadillo (none of which exist in this container).
//
// TEST INFRASTRUCTURE ONLY.  It implements exactly the subset those headers use:
// dense column-major mat/vec/rowvec, cube, field<T>, row/col/slice views, + - * / with
// scalars and matrices, t(), dot, accu, zeros/ones/eye/diagmat, pinv, inv, mvnrnd, and
// the R:: r*/d* functions.  Every random draw is popped from a "tape" the test injects
// (shim::tape()); when the tape is empty a std::mt19937_64 stream is used instead, so the
// reference's statistical-recovery tests can also be run.
//
// What is NOT the reference here: the container arithmetic (matrix products, pinv via
// Jacobi, inv via Gauss-Jordan, mvnrnd = mean + chol_lower(C) z as documented for
// arma::mvnrnd) and R's generators.  The update LOGIC is the reference's, compiled from
// where it lies.
#ifndef BFMMM_SHIM_RCPPARMADILLO_H
#define BFMMM_SHIM_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <deque>
#include <iostream>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace shim {
struct Tape {
  std::deque<double> q;
  std::mt19937_64 rng{12345};
  long popped = 0;
  bool has() const { return !q.empty(); }
  double pop() { double v = q.front(); q.pop_front(); popped++; return v; }
};
inline Tape& tape() { static Tape t; return t; }
inline double norm_rand() {
  Tape& t = tape();
  if (t.has()) return t.pop();
  return std::normal_distribution<double>(0.0, 1.0)(t.rng);
}
inline double unif_rand() {
  Tape& t = tape();
  if (t.has()) return t.pop();
  return std::uniform_real_distribution<double>(0.0, 1.0)(t.rng);
}
inline double gamma_rand(double shape) {
  Tape& t = tape();
  if (t.has()) return t.pop();
  return std::gamma_distribution<double>(shape, 1.0)(t.rng);
}
}  // namespace shim

namespace arma {
typedef unsigned long long uword;
namespace fill {
struct zeros_t {}; struct ones_t {}; struct randn_t {};
static const zeros_t zeros = zeros_t();
static const ones_t ones = ones_t();
static const randn_t randn = randn_t();
}
namespace datum { static const double pi = 3.14159265358979323846; }

class Mat;

// non-owning rectangular view
class subview {
 public:
  double* base; uword ld; uword n_rows, n_cols, n_elem;
  subview(double* b, uword ld_, uword r, uword c) : base(b), ld(ld_), n_rows(r), n_cols(c), n_elem(r * c) {}
  double& at(uword r, uword c) const { return base[c * ld + r]; }
  double& operator()(uword i) const { return n_rows == 1 ? base[i * ld] : base[i]; }
  double& operator()(uword r, uword c) const { return at(r, c); }
  inline Mat t() const;
  inline subview& operator=(const Mat& m);
  subview& operator=(const subview& o) {
    std::vector<double> tmp(o.n_elem);
    for (uword c = 0; c < o.n_cols; c++) for (uword r = 0; r < o.n_rows; r++) tmp[c * o.n_rows + r] = o.at(r, c);
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) at(r, c) = tmp[c * n_rows + r];
    return *this;
  }
  subview(const subview&) = default;
};

class Mat {
 public:
  uword n_rows, n_cols, n_elem;
  std::vector<double> mem;
  Mat() : n_rows(0), n_cols(0), n_elem(0) {}
  Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, 0.0) {}
  Mat(uword r, uword c, fill::zeros_t) : Mat(r, c) {}
  Mat(uword r, uword c, fill::ones_t) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, 1.0) {}
  Mat(uword r, uword c, fill::randn_t) : Mat(r, c) { for (auto& x : mem) x = shim::norm_rand(); }
  Mat(const subview& s) : n_rows(s.n_rows), n_cols(s.n_cols), n_elem(s.n_elem), mem(s.n_elem) {
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) mem[c * n_rows + r] = s.at(r, c);
  }
  double* memptr() { return mem.data(); }
  const double* memptr() const { return mem.data(); }
  double& operator()(uword i) { return mem[i]; }
  const double& operator()(uword i) const { return mem[i]; }
  double& operator[](uword i) { return mem[i]; }
  const double& operator[](uword i) const { return mem[i]; }
  double& operator()(uword r, uword c) { return mem[c * n_rows + r]; }
  const double& operator()(uword r, uword c) const { return mem[c * n_rows + r]; }
  subview row(uword r) const { return subview(const_cast<double*>(mem.data()) + r, n_rows, 1, n_cols); }
  subview col(uword c) const { return subview(const_cast<double*>(mem.data()) + c * n_rows, n_rows, n_rows, 1); }
  Mat t() const {
    Mat o(n_cols, n_rows);
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) o.mem[r * n_cols + c] = mem[c * n_rows + r];
    return o;
  }
  Mat& zeros() { std::fill(mem.begin(), mem.end(), 0.0); return *this; }
  Mat& ones() { std::fill(mem.begin(), mem.end(), 1.0); return *this; }
  double min() const { return *std::min_element(mem.begin(), mem.end()); }
  struct diag_view {
    Mat& m;
    std::vector<double> vals() const { std::vector<double> v(m.n_rows); for (uword i = 0; i < m.n_rows; i++) v[i] = m(i, i); return v; }
    std::vector<double> operator+(double s) const { auto v = vals(); for (auto& x : v) x += s; return v; }
    diag_view& operator=(const std::vector<double>& v) { for (uword i = 0; i < m.n_rows; i++) m(i, i) = v[i]; return *this; }
  };
  diag_view diag() { return diag_view{*this}; }
};

inline Mat subview::t() const {
  Mat o(n_cols, n_rows);
  for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) o.mem[r * n_cols + c] = at(r, c);
  return o;
}
inline subview& subview::operator=(const Mat& m) {
  for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) at(r, c) = m.mem[c * n_rows + r];
  return *this;
}

class Col : public Mat {
 public:
  Col() : Mat() {}
  explicit Col(uword n) : Mat(n, 1) {}
  Col(uword n, fill::zeros_t) : Mat(n, 1) {}
  Col(uword n, fill::ones_t) : Mat(n, 1, fill::ones) {}
  Col(uword n, fill::randn_t) : Mat(n, 1, fill::randn) {}
  Col(const Mat& m) : Mat(m) {}
  Col(const subview& s) : Mat(s) {}
  Col(std::initializer_list<double> l) : Mat(l.size(), 1) { std::copy(l.begin(), l.end(), mem.begin()); }
};
class Row : public Mat {
 public:
  Row() : Mat() {}
  explicit Row(uword n) : Mat(1, n) {}
  Row(uword n, fill::zeros_t) : Mat(1, n) {}
  Row(const Mat& m) : Mat(m) {}
  Row(const subview& s) : Mat(s) {}
};
typedef Mat mat;
typedef Col vec;
typedef Row rowvec;

class Cube {
 public:
  uword n_rows, n_cols, n_slices, n_elem;
  std::vector<Mat> sl;
  Cube() : n_rows(0), n_cols(0), n_slices(0), n_elem(0) {}
  Cube(uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s), sl(s, Mat(r, c)) {}
  Cube(uword r, uword c, uword s, fill::zeros_t) : Cube(r, c, s) {}
  Cube(uword r, uword c, uword s, fill::ones_t) : n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s), sl(s, Mat(r, c, fill::ones)) {}
  Cube(uword r, uword c, uword s, fill::randn_t) : Cube(r, c, s) { for (auto& m : sl) for (auto& x : m.mem) x = shim::norm_rand(); }
  Mat& slice(uword s) { return sl[s]; }
  const Mat& slice(uword s) const { return sl[s]; }
  double& operator()(uword r, uword c, uword s) { return sl[s](r, c); }
  const double& operator()(uword r, uword c, uword s) const { return sl[s](r, c); }
};
typedef Cube cube;

template <typename T>
class field {
 public:
  uword n_rows, n_cols, n_elem;
  std::vector<T> el;
  field() : n_rows(0), n_cols(0), n_elem(0) {}
  field(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), el(r * c) {}
  T& operator()(uword r, uword c) { return el[c * n_rows + r]; }
  const T& operator()(uword r, uword c) const { return el[c * n_rows + r]; }
};

// ---- arithmetic (eager) ----
inline Mat operator+(const Mat& a, const Mat& b) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = a.mem[i] + b.mem[i]; return o; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = a.mem[i] - b.mem[i]; return o; }
inline Mat operator*(double s, const Mat& a) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = s * a.mem[i]; return o; }
inline Mat operator*(const Mat& a, double s) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = a.mem[i] * s; return o; }
inline Mat operator/(const Mat& a, double s) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = a.mem[i] / s; return o; }
inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.n_cols != b.n_rows) throw std::logic_error("shim: matrix multiplication: incompatible dimensions");
  Mat o(a.n_rows, b.n_cols);
  for (uword c = 0; c < b.n_cols; c++)
    for (uword k = 0; k < a.n_cols; k++) {
      double bkc = b.mem[c * b.n_rows + k];
      for (uword r = 0; r < a.n_rows; r++) o.mem[c * a.n_rows + r] += a.mem[k * a.n_rows + r] * bkc;
    }
  return o;
}
// subview operands (exact-match overloads keep the hot dot products allocation-free)
inline double dot(const subview& a, const subview& b) {
  double s = 0; uword n = a.n_elem;
  if (a.n_rows == 1 && b.n_rows == 1) { for (uword i = 0; i < n; i++) s += a.base[i * a.ld] * b.base[i * b.ld]; return s; }
  for (uword i = 0; i < n; i++) s += a(i) * b(i);
  return s;
}
inline double dot(const Mat& a, const Mat& b) { double s = 0; for (uword i = 0; i < a.n_elem; i++) s += a.mem[i] * b.mem[i]; return s; }
inline double dot(const Mat& a, const subview& b) { double s = 0; for (uword i = 0; i < a.n_elem; i++) s += a.mem[i] * b(i); return s; }
inline double dot(const subview& a, const Mat& b) { double s = 0; for (uword i = 0; i < b.n_elem; i++) s += a(i) * b.mem[i]; return s; }
inline double accu(const Mat& a) { double s = 0; for (double x : a.mem) s += x; return s; }
inline double accu(const subview& a) { double s = 0; for (uword c = 0; c < a.n_cols; c++) for (uword r = 0; r < a.n_rows; r++) s += a.at(r, c); return s; }

inline Col zeros(uword n) { return Col(n); }
inline Mat zeros(uword r, uword c) { return Mat(r, c); }
inline Cube zeros(uword r, uword c, uword s) { return Cube(r, c, s); }
inline Col ones(uword n) { return Col(n, fill::ones); }
inline Mat ones(uword r, uword c) { return Mat(r, c, fill::ones); }
inline Mat eye(uword r, uword c) { Mat o(r, c); for (uword i = 0; i < std::min(r, c); i++) o(i, i) = 1.0; return o; }
inline Mat diagmat(const Mat& v) {
  if (v.n_rows != 1 && v.n_cols != 1) { Mat o(v.n_rows, v.n_cols); for (uword i = 0; i < std::min(v.n_rows, v.n_cols); i++) o(i, i) = v(i, i); return o; }
  Mat o(v.n_elem, v.n_elem); for (uword i = 0; i < v.n_elem; i++) o(i, i) = v.mem[i]; return o;
}
inline Mat floor(const Mat& a) { Mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) o.mem[i] = std::floor(a.mem[i]); return o; }

// ---- decompositions (documented semantics; implementations are the shim's own) ----
inline bool chol_lower(const Mat& A, Mat& L) {
  uword n = A.n_rows; L = Mat(n, n);
  for (uword j = 0; j < n; j++) {
    double s = A(j, j);
    for (uword k = 0; k < j; k++) s -= L(j, k) * L(j, k);
    if (!(s > 0)) return false;
    double d = std::sqrt(s); L(j, j) = d;
    for (uword i = j + 1; i < n; i++) {
      double t = A(i, j);
      for (uword k = 0; k < j; k++) t -= L(i, k) * L(j, k);
      L(i, j) = t / d;
    }
  }
  return true;
}
inline bool inv(Mat& out, const Mat& A) {
  uword n = A.n_rows; Mat a = A; Mat b = eye(n, n);
  for (uword c = 0; c < n; c++) {
    uword piv = c; double best = std::fabs(a(c, c));
    for (uword r = c + 1; r < n; r++) if (std::fabs(a(r, c)) > best) { best = std::fabs(a(r, c)); piv = r; }
    if (best == 0.0) throw std::runtime_error("shim: inv(): matrix is singular");
    if (piv != c) for (uword j = 0; j < n; j++) { std::swap(a(c, j), a(piv, j)); std::swap(b(c, j), b(piv, j)); }
    double iv = 1.0 / a(c, c);
    for (uword j = 0; j < n; j++) { a(c, j) *= iv; b(c, j) *= iv; }
    for (uword r = 0; r < n; r++) {
      if (r == c) continue; double f = a(r, c); if (f == 0.0) continue;
      for (uword j = 0; j < n; j++) { a(r, j) -= f * a(c, j); b(r, j) -= f * b(c, j); }
    }
  }
  out = b; return true;
}
inline Mat inv(const Mat& A) { Mat o; inv(o, A); return o; }
inline Mat inv_sympd(const Mat& A) { return inv(A); }
inline double log_det_sympd(const Mat& A) { Mat L; if (!chol_lower(A, L)) throw std::runtime_error("shim: log_det_sympd"); double s = 0; for (uword i = 0; i < A.n_rows; i++) s += std::log(L(i, i)); return 2 * s; }
// pinv of a symmetric matrix by cyclic Jacobi; tolerance n*max|lambda|*eps (Armadillo default)
inline Mat pinv(const Mat& A) {
  uword n = A.n_rows; Mat a = A; Mat V = eye(n, n);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0; for (uword p = 0; p < n; p++) for (uword q = p + 1; q < n; q++) off += a(p, q) * a(p, q);
    if (off < 1e-300) break;
    for (uword p = 0; p < n; p++) for (uword q = p + 1; q < n; q++) {
      double apq = a(p, q); if (apq == 0.0) continue;
      double theta = (a(q, q) - a(p, p)) / (2 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
      double c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (uword k = 0; k < n; k++) { double akp = a(k, p), akq = a(k, q); a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq; }
      for (uword k = 0; k < n; k++) { double apk = a(p, k), aqk = a(q, k); a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk; }
      for (uword k = 0; k < n; k++) { double vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
    }
  }
  double lmax = 0; for (uword i = 0; i < n; i++) lmax = std::max(lmax, std::fabs(a(i, i)));
  double tol = n * lmax * std::numeric_limits<double>::epsilon();
  Mat o(n, n);
  for (uword e = 0; e < n; e++) {
    double lam = a(e, e); if (std::fabs(lam) <= tol) continue;
    for (uword j = 0; j < n; j++) { double vj = V(j, e) / lam; for (uword i = 0; i < n; i++) o(i, j) += V(i, e) * vj; }
  }
  return o;
}
// arma::mvnrnd(M, C): M + chol(C, "lower") * randn
inline Col mvnrnd(const Mat& M, const Mat& C) {
  Mat L; if (!chol_lower(C, L)) throw std::runtime_error("shim: mvnrnd(): given covariance matrix is not symmetric positive semi-definite");
  uword n = M.n_elem; std::vector<double> z(n); for (auto& x : z) x = shim::norm_rand();
  Col o(n);
  for (uword i = 0; i < n; i++) { double d = 0; for (uword j = 0; j <= i; j++) d += L(i, j) * z[j]; o.mem[i] = M.mem[i] + d; }
  return o;
}
inline Cube randn(uword r, uword c, uword s) { return Cube(r, c, s, fill::randn); }
}  // namespace arma

// ---- R's Rmath entry points used by the reference ----
namespace R {
inline double rnorm(double mu, double sd) { return mu + sd * shim::norm_rand(); }
inline double runif(double a, double b) { return a + (b - a) * shim::unif_rand(); }
inline double rgamma(double shape, double scale) { return scale * shim::gamma_rand(shape); }
inline double rbeta(double a, double b) { double x = shim::gamma_rand(a), y = shim::gamma_rand(b); return x / (x + y); }
inline double rbinom(double n, double p) { double c = 0; for (int i = 0; i < (int)n; i++) c += shim::unif_rand() < p; return c; }
inline double dnorm(double x, double mu, double sd, int lg) {
  const double M_LN_SQRT_2PI_ = 0.918938533204672741780329736406;
  double z = (x - mu) / sd;
  double l = -(M_LN_SQRT_2PI_ + 0.5 * z * z + std::log(sd));
  return lg ? l : std::exp(l);
}
inline double pnorm(double x, double mu, double sd, int lower, int lg) {
  double p = 0.5 * std::erfc(-(x - mu) / (sd * std::sqrt(2.0)));
  if (!lower) p = 1 - p;
  return lg ? std::log(p) : p;
}
}  // namespace R

namespace Rcpp {
// progress output of the reference is discarded (a stream without a buffer: every insertion is a no-op)
static std::ostream Rcout_null(nullptr);
static std::ostream& Rcout = Rcout_null;
inline void checkUserInterrupt() {}
}

#endif
