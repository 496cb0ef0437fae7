// bfmmm_oracle.cpp -- CPU restatement of BayesFMMM's per-iteration updates.
// TEST INFRASTRUCTURE ONLY (see bfmmm_oracle.h).  Plain C++17, no dependencies.
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference/inst/include/BayesFMMM/).  The four reference variants of each
// update (functional, multivariate, covariate-adjusted functional, covariate-adjusted
// multivariate) are the same loop with (a) B_i = I for the multivariate model and
// (b) nu_k -> nu_k + eta_k x_i, phi_km -> phi_km + xi_km x_i for the covariate-adjusted
// one; the restatement is written once over those two switches.
#include "bfmmm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct View {
  const orc_data* d;
  const orc_state* s;
  int n, K, P, M, D;
  bool ident;
  // mutable working copies of the global blocks (Armadillo layouts)
  std::vector<double> nu, Phi, eta, xi;

  View(const orc_data* d_, const orc_state* s_) : d(d_), s(s_) {
    n = d->n; K = d->K; P = d->P; M = d->M; D = d->D; ident = d->identity_basis != 0;
    nu.assign(s->nu, s->nu + (size_t)K * P);
    if (M > 0 && s->Phi) Phi.assign(s->Phi, s->Phi + (size_t)K * P * M);
    if (D > 0) {
      eta.assign(s->eta, s->eta + (size_t)P * D * K);
      if (M > 0) xi.assign(s->xi, s->xi + (size_t)K * P * D * M);
    }
  }
  int64_t npts(int i) const { return ident ? P : d->off[i + 1] - d->off[i]; }
  double y(int i, int64_t l) const { return ident ? d->y[(size_t)l * n + i] : d->y[d->off[i] + l]; }
  const double* Brow(int i, int64_t l) const { return d->B + (size_t)(d->off[i] + l) * P; }
  double Z(int i, int k) const { return s->Z[(size_t)k * n + i]; }
  double chi(int i, int m) const { return s->chi[(size_t)m * n + i]; }
  double X(int i, int dd) const { return d->X[(size_t)dd * n + i]; }
  double& nu_(int k, int p) { return nu[(size_t)p * K + k]; }
  double& Phi_(int k, int p, int m) { return Phi[((size_t)m * P + p) * K + k]; }
  double& eta_(int p, int dd, int k) { return eta[((size_t)k * D + dd) * P + p]; }
  double& xi_(int k, int p, int dd, int m) { return xi[(size_t)k * P * D * M + ((size_t)m * D + dd) * P + p]; }
};

// Per-function effective coefficients nu_k + eta_k x_i and phi_km + xi_km x_i evaluated at
// point l: an[k] = B_l . (nu_k + eta_k x_i), fn[k*M+m] = B_l . (phi_km + xi_km x_i).
// (lpdf_zCovariateAdj UpdateMixedMembership.h:525-536 evaluates the same two dots and adds.)
struct PointEval {
  View& v;
  std::vector<double> an, fn, ex, xx;   // ex: eta_k x_i (P), xx: xi_km x_i (P)
  std::vector<double> Ak, Fkm;          // effective coefficient vectors of function i
  explicit PointEval(View& v_) : v(v_), an(v_.K), fn((size_t)v_.K * std::max(v_.M, 1)),
        Ak((size_t)v_.K * v_.P), Fkm((size_t)v_.K * std::max(v_.M, 1) * v_.P) {}
  void set_function(int i) {
    const int K = v.K, P = v.P, M = v.M, D = v.D;
    for (int k = 0; k < K; k++)
      for (int p = 0; p < P; p++) {
        double a = 0;
        for (int dd = 0; dd < D; dd++) a += v.eta_(p, dd, k) * v.X(i, dd);
        Ak[(size_t)k * P + p] = a;     // covariate part only; nu added per point like the reference
        for (int m = 0; m < M; m++) {
          double f = 0;
          for (int dd = 0; dd < D; dd++) f += v.xi_(k, p, dd, m) * v.X(i, dd);
          Fkm[((size_t)k * M + m) * P + p] = f;
        }
      }
  }
  void eval(int i, int64_t l) {
    const int K = v.K, P = v.P, M = v.M, D = v.D;
    if (v.ident) {
      for (int k = 0; k < K; k++) {
        an[k] = v.nu_(k, (int)l) + (D ? Ak[(size_t)k * P + l] : 0.0);
        for (int m = 0; m < M; m++)
          fn[(size_t)k * M + m] = v.Phi_(k, (int)l, m) + (D ? Fkm[((size_t)k * M + m) * P + l] : 0.0);
      }
      return;
    }
    const double* b = v.Brow(i, l);
    for (int k = 0; k < K; k++) {
      double dn = 0, de = 0;
      for (int p = 0; p < P; p++) dn += v.nu_(k, p) * b[p];
      if (D) for (int p = 0; p < P; p++) de += Ak[(size_t)k * P + p] * b[p];
      an[k] = D ? dn + de : dn;
      for (int m = 0; m < M; m++) {
        double dp = 0, dx = 0;
        for (int p = 0; p < P; p++) dp += v.Phi_(k, p, m) * b[p];
        if (D) for (int p = 0; p < P; p++) dx += Fkm[((size_t)k * M + m) * P + p] * b[p];
        fn[(size_t)k * M + m] = D ? dp + dx : dp;
      }
    }
  }
};

// ---------------------------------------------------------------- small dense linear algebra
// column-major n x n
bool chol_lower(int n, const double* A, double* L) {
  std::fill(L, L + (size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) s -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
    if (!(s > 0)) return false;
    double ljj = std::sqrt(s);
    L[(size_t)j * n + j] = ljj;
    for (int i = j + 1; i < n; i++) {
      double t = A[(size_t)j * n + i];
      for (int k = 0; k < j; k++) t -= L[(size_t)k * n + i] * L[(size_t)k * n + j];
      L[(size_t)j * n + i] = t / ljj;
    }
  }
  return true;
}

// Gauss-Jordan inverse with partial pivoting (stand-in for arma::inv, UpdatePhi.h:79)
bool inv_general(int n, const double* A, double* Ainv) {
  std::vector<double> a(A, A + (size_t)n * n);
  std::vector<double> b((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) b[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c; double best = std::fabs(a[(size_t)c * n + c]);
    for (int r = c + 1; r < n; r++)
      if (std::fabs(a[(size_t)c * n + r]) > best) { best = std::fabs(a[(size_t)c * n + r]); piv = r; }
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; j++) {
        std::swap(a[(size_t)j * n + c], a[(size_t)j * n + piv]);
        std::swap(b[(size_t)j * n + c], b[(size_t)j * n + piv]);
      }
    double inv = 1.0 / a[(size_t)c * n + c];
    for (int j = 0; j < n; j++) { a[(size_t)j * n + c] *= inv; b[(size_t)j * n + c] *= inv; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[(size_t)c * n + r];
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) {
        a[(size_t)j * n + r] -= f * a[(size_t)j * n + c];
        b[(size_t)j * n + r] -= f * b[(size_t)j * n + c];
      }
    }
  }
  std::copy(b.begin(), b.end(), Ainv);
  return true;
}

// Moore-Penrose inverse of a symmetric matrix by cyclic Jacobi (stand-in for arma::pinv,
// UpdateNu.h:67; tolerance = n * max|lambda| * eps as in Armadillo's default).
bool pinv_sym(int n, const double* A, double* Ainv) {
  std::vector<double> a(A, A + (size_t)n * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)q * n + p] * a[(size_t)q * n + p];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[(size_t)q * n + p];
        if (apq == 0.0) continue;
        double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        double theta = (aqq - app) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), sn = t * c;
        for (int k = 0; k < n; k++) {   // rotate columns p,q
          double akp = a[(size_t)p * n + k], akq = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * akp - sn * akq;
          a[(size_t)q * n + k] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {   // rotate rows p,q
          double apk = a[(size_t)k * n + p], aqk = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * apk - sn * aqk;
          a[(size_t)k * n + q] = sn * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = V[(size_t)p * n + k], vkq = V[(size_t)q * n + k];
          V[(size_t)p * n + k] = c * vkp - sn * vkq;
          V[(size_t)q * n + k] = sn * vkp + c * vkq;
        }
      }
  }
  double lmax = 0;
  for (int i = 0; i < n; i++) lmax = std::max(lmax, std::fabs(a[(size_t)i * n + i]));
  double tol = n * lmax * 2.220446049250313e-16;
  std::fill(Ainv, Ainv + (size_t)n * n, 0.0);
  for (int e = 0; e < n; e++) {
    double lam = a[(size_t)e * n + e];
    if (std::fabs(lam) <= tol) continue;
    double il = 1.0 / lam;
    for (int j = 0; j < n; j++) {
      double vj = V[(size_t)e * n + j] * il;
      for (int i = 0; i < n; i++) Ainv[(size_t)j * n + i] += V[(size_t)e * n + i] * vj;
    }
  }
  return true;
}

// arma::mvnrnd(mean, C): mean + chol_lower(C) * z   (UpdateNu.h:69, UpdatePhi.h:82)
bool mvn_draw(int n, const double* C, const double* rhs, const double* z, double* out) {
  std::vector<double> L((size_t)n * n);
  if (!chol_lower(n, C, L.data())) return false;
  for (int i = 0; i < n; i++) {
    double mean = 0, dev = 0;
    for (int j = 0; j < n; j++) mean += C[(size_t)j * n + i] * rhs[j];
    for (int j = 0; j <= i; j++) dev += L[(size_t)j * n + i] * z[j];
    out[i] = mean + dev;
  }
  return true;
}

// ---------------------------------------------------------------- generic Gaussian block update
enum BlockType { B_NU = 0, B_ETA = 1, B_PHI = 2, B_XI = 3 };
struct Block { int type, k, m, d; };

struct BlockEngine {
  View& v;
  std::vector<std::vector<double>> Gi;   // cached sum_l B_l B_l' per function (functional only)
  explicit BlockEngine(View& v_) : v(v_) {}

  double weight(const Block& b, int i) const {
    double w = v.Z(i, b.k);
    if (b.type == B_PHI || b.type == B_XI) w *= v.chi(i, b.m);
    if (b.type == B_ETA || b.type == B_XI) w *= v.X(i, b.d);
    return w;
  }
  double coef(const Block& b, int p) {
    switch (b.type) {
      case B_NU: return v.nu_(b.k, p);
      case B_ETA: return v.eta_(p, b.d, b.k);
      case B_PHI: return v.Phi_(b.k, p, b.m);
      default: return v.xi_(b.k, p, b.d, b.m);
    }
  }
  void set_coef(const Block& b, int p, double x) {
    switch (b.type) {
      case B_NU: v.nu_(b.k, p) = x; break;
      case B_ETA: v.eta_(p, b.d, b.k) = x; break;
      case B_PHI: v.Phi_(b.k, p, b.m) = x; break;
      default: v.xi_(b.k, p, b.d, b.m) = x; break;
    }
  }
  std::vector<Block> all_blocks() const {
    std::vector<Block> bs;
    for (int k = 0; k < v.K; k++) {
      bs.push_back({B_NU, k, 0, 0});
      for (int d = 0; d < v.D; d++) bs.push_back({B_ETA, k, 0, d});
      for (int m = 0; m < v.M; m++) {
        bs.push_back({B_PHI, k, m, 0});
        for (int d = 0; d < v.D; d++) bs.push_back({B_XI, k, m, d});
      }
    }
    return bs;
  }

  // Accumulate precision M1 (P x P) and rhs m1 (P) of the target block exactly as the
  // reference loops do: M1 += w_ia^2 B_l B_l', m1 += w_ia B_l (y_l - sum_{b != a} w_ib B_l.c_b)
  // (UpdateNu.h:42-63, UpdatePhi.h:44-71, UpdateEta.h:51-81, UpdateXi.h:51-72).
  void accumulate(const Block& tgt, std::vector<double>& M1, std::vector<double>& m1) {
    const int P = v.P, n = v.n;
    std::fill(M1.begin(), M1.end(), 0.0);
    std::fill(m1.begin(), m1.end(), 0.0);
    std::vector<Block> bs = all_blocks();
    std::vector<double> theta(P);
    std::vector<int> nz;
    for (int i = 0; i < n; i++) {
      if (v.Z(i, tgt.k) == 0) continue;
      double wa = weight(tgt, i);
      // coefficient vector of everything except the target block
      std::fill(theta.begin(), theta.end(), 0.0);
      for (const Block& b : bs) {
        if (b.type == tgt.type && b.k == tgt.k && b.m == tgt.m && b.d == tgt.d) continue;
        double wb = weight(b, i);
        if (wb == 0) continue;
        for (int p = 0; p < P; p++) theta[p] += wb * coef(b, p);
      }
      if (v.ident) {
        for (int p = 0; p < P; p++) {
          M1[(size_t)p * P + p] += wa * wa;
          m1[p] += wa * (v.y(i, p) - theta[p]);
        }
        continue;
      }
      const int64_t T = v.npts(i);
      for (int64_t l = 0; l < T; l++) {
        const double* b = v.Brow(i, l);
        double mean = 0;
        for (int p = 0; p < P; p++) mean += theta[p] * b[p];
        double ph = v.y(i, l) - mean;
        double s2 = wa * wa;
        // outer product B_l B_l' restricted to the non-zero basis values of this point (the skipped
        // terms are exact zeros: a tensor-product cubic basis has 16 of P = 400)
        nz.clear();
        for (int c = 0; c < P; c++) if (b[c] != 0.0) nz.push_back(c);
        for (int c : nz) {
          double sc = s2 * b[c];
          for (int r : nz) M1[(size_t)c * P + r] += sc * b[r];
        }
        double wp = wa * ph;
        for (int p = 0; p < P; p++) m1[p] += wp * b[p];
      }
    }
  }
};

int draw_block(BlockEngine& be, const Block& tgt, const std::vector<double>& prior /* P x P */,
               bool use_pinv, double beta, double sigma_sq, const double* z) {
  const int P = be.v.P;
  std::vector<double> M1((size_t)P * P), m1(P), C((size_t)P * P), out(P);
  be.accumulate(tgt, M1, m1);
  for (int p = 0; p < P; p++) m1[p] = (m1[p] * beta) / sigma_sq;
  for (size_t e = 0; e < M1.size(); e++) M1[e] = (M1[e] * beta) / sigma_sq + prior[e];
  if (use_pinv) {
    if (!pinv_sym(P, M1.data(), C.data())) return 1;
    for (int c = 0; c < P; c++)
      for (int r = c + 1; r < P; r++) {
        double s = (C[(size_t)c * P + r] + C[(size_t)r * P + c]) / 2;
        C[(size_t)c * P + r] = s; C[(size_t)r * P + c] = s;
      }
  } else {
    if (!inv_general(P, M1.data(), C.data())) return 1;
  }
  if (!mvn_draw(P, C.data(), m1.data(), z, out.data())) return 2;
  for (int p = 0; p < P; p++) be.set_coef(tgt, p, out[p]);
  return 0;
}

}  // namespace

// ================================================================= basis construction
// Clamped B-spline design matrix with intercept, the 4-argument form
// splines2::BSpline(t, internal_knots, degree, boundary_knots).basis(true) used at
// BFMMM.h:1188-1196 (splines2 is not vendored; behaviour pinned by Tensor_BSpline.txt).
extern "C" int orc_bspline_basis(const double* t, int64_t n, const double* iknots, int n_ik,
                                 int degree, double b_lo, double b_hi, double* B) {
  const int P = n_ik + degree + 1;
  const int nk = n_ik + 2 * (degree + 1);
  std::vector<double> kn(nk);
  for (int i = 0; i <= degree; i++) { kn[i] = b_lo; kn[nk - 1 - i] = b_hi; }
  for (int i = 0; i < n_ik; i++) kn[degree + 1 + i] = iknots[i];
  std::vector<double> h(degree + 1), hh(degree + 1);
  for (int64_t r = 0; r < n; r++) {
    double x = t[r];
    double* row = B + (size_t)r * P;
    std::fill(row, row + P, 0.0);
    if (x < b_lo || x > b_hi) continue;
    int ell = degree;                       // knot span: kn[ell] <= x < kn[ell+1]
    while (ell < nk - degree - 2 && x >= kn[ell + 1]) ell++;
    h[0] = 1.0;
    for (int j = 1; j <= degree; j++) {
      for (int q = 0; q < j; q++) hh[q] = h[q];
      h[0] = 0.0;
      for (int q = 1; q <= j; q++) {
        int ind = ell + q;
        double xb = kn[ind], xa = kn[ind - j];
        if (xb == xa) { h[q] = 0.0; continue; }
        double w = hh[q - 1] / (xb - xa);
        h[q - 1] += w * (xb - x);
        h[q] = w * (x - xa);
      }
    }
    for (int q = 0; q <= degree; q++) row[ell - degree + q] = h[q];
  }
  return 0;
}

// TensorBSpline, BSplines.h:18-62: column i <-> multi-index with dimension 0 slowest.
extern "C" int orc_tensor_bspline(const double* t, int64_t n, int dim, const int* degree,
                                  const double* boundary, const double* iknots_concat,
                                  const int* n_ik, double* B) {
  std::vector<int> Pd(dim);
  int P = 1;
  for (int l = 0; l < dim; l++) { Pd[l] = n_ik[l] + degree[l] + 1; P *= Pd[l]; }
  std::vector<std::vector<double>> Bd(dim);
  const double* ik = iknots_concat;
  for (int l = 0; l < dim; l++) {
    Bd[l].resize((size_t)n * Pd[l]);
    orc_bspline_basis(t + (size_t)l * n, n, ik, n_ik[l], degree[l], boundary[2 * l], boundary[2 * l + 1],
                      Bd[l].data());
    ik += n_ik[l];
  }
  std::vector<int> idx(dim);
  for (int c = 0; c < P; c++) {
    int rem = c;
    for (int l = dim - 1; l >= 0; l--) { idx[l] = rem % Pd[l]; rem /= Pd[l]; }
    for (int64_t r = 0; r < n; r++) {
      double val = 1.0;                      // B = ones, then multiplied dimension by dimension (:35,49-51)
      for (int l = 0; l < dim; l++) val = val * Bd[l][(size_t)r * Pd[l] + idx[l]];
      B[(size_t)r * P + c] = val;
    }
  }
  return 0;
}

// GetP, BSplines.h:70-120: P = C'C with one first-difference row per neighbouring index pair.
extern "C" int orc_getP(int dim, const int* degree, const int* n_ik, double* Pmat) {
  std::vector<int> Pd(dim);
  int P = 1;
  for (int l = 0; l < dim; l++) { Pd[l] = n_ik[l] + degree[l] + 1; P *= Pd[l]; }
  std::vector<std::vector<int>> index(P, std::vector<int>(dim));
  for (int c = 0; c < P; c++) {
    int rem = c;
    for (int l = dim - 1; l >= 0; l--) { index[c][l] = rem % Pd[l]; rem /= Pd[l]; }
  }
  std::fill(Pmat, Pmat + (size_t)P * P, 0.0);
  for (int i = 0; i < P; i++)
    for (int j = i; j < P; j++) {
      int diff = 0, adiff = 0;
      for (int l = 0; l < dim; l++) { diff += index[j][l] - index[i][l]; adiff += std::abs(index[j][l] - index[i][l]); }
      if (diff == 1 && adiff == 1) {         // constraint row e_i - e_j (:103-107)
        Pmat[(size_t)i * P + i] += 1; Pmat[(size_t)j * P + j] += 1;
        Pmat[(size_t)j * P + i] -= 1; Pmat[(size_t)i * P + j] -= 1;
      }
    }
  return 0;
}

// first-order random-walk penalty, BFMMM.h:1198-1208
extern "C" void orc_pmat_rw1(int P, double* Pmat) {
  std::fill(Pmat, Pmat + (size_t)P * P, 0.0);
  for (int j = 0; j < P; j++) {
    Pmat[0] = 1;
    if (j > 0) {
      Pmat[(size_t)j * P + j] = 2;
      Pmat[(size_t)j * P + (j - 1)] = -1;
      Pmat[(size_t)(j - 1) * P + j] = -1;
    }
    Pmat[(size_t)(P - 1) * P + (P - 1)] = 1;
  }
}

// ================================================================= Z Metropolis step
// updateZ_PM UpdateMixedMembership.h:131-185 (lpdf_z :20-50, Z_proposal_density :102-113,
// rdirichlet Distributions.h:22-45, calc_lB :51-61); MV :273-300,358-411; covariate-adjusted
// :506-541,615-674; MV covariate-adjusted :768-798,866-922; tempered twins multiply the
// squared-error term by beta (:91).
extern "C" int orc_update_z(const orc_data* d, const orc_state* s, const double* pi, double alpha3,
                            double a_Z_PM, double beta, const double* gam, const double* u,
                            double* Z_out, double* acc_out, int32_t* accepted) {
  View v(d, s);
  PointEval pe(v);
  const int n = v.n, K = v.K, M = v.M;
  std::vector<double> zc(K), zp(K), al(K);
  for (int i = 0; i < n; i++) {
    if (v.D) pe.set_function(i);
    for (int k = 0; k < K; k++) zc[k] = v.Z(i, k);
    // proposal: rdirichlet(a_Z_PM * z) with non-positive concentrations replaced by 10
    double sum = 0;
    for (int k = 0; k < K; k++) { double g = gam[(size_t)k * n + i]; zp[k] = g; sum += g; }
    for (int k = 0; k < K; k++) zp[k] = zp[k] / sum;
    double lp_old = 0, lp_new = 0;
    for (int k = 0; k < K; k++) {
      lp_old += (alpha3 * pi[k] - 1) * std::log(zc[k]);
      lp_new += (alpha3 * pi[k] - 1) * std::log(zp[k]);
    }
    const int64_t T = v.npts(i);
    for (int64_t l = 0; l < T; l++) {
      pe.eval(i, l);
      double mo = 0, mn = 0;
      for (int k = 0; k < K; k++) {
        mo += zc[k] * pe.an[k];
        mn += zp[k] * pe.an[k];
        for (int m = 0; m < M; m++) {
          mo += zc[k] * v.chi(i, m) * pe.fn[(size_t)k * M + m];
          mn += zp[k] * v.chi(i, m) * pe.fn[(size_t)k * M + m];
        }
      }
      double yl = v.y(i, l);
      lp_old -= beta * (std::pow(yl - mo, 2.0) / (2 * s->sigma_sq));
      lp_new -= beta * (std::pow(yl - mn, 2.0) / (2 * s->sigma_sq));
    }
    // q(new | a*old) and q(old | a*new)
    auto prop_density = [&](const std::vector<double>& x, const std::vector<double>& from) {
      double dens = 0, lB = 0, tot = 0;
      for (int k = 0; k < K; k++) { al[k] = a_Z_PM * from[k]; }
      for (int k = 0; k < K; k++) dens += (al[k] - 1) * std::log(x[k]);
      for (int k = 0; k < K; k++) { lB += std::lgamma(al[k]); tot += al[k]; }
      lB -= std::lgamma(tot);
      return dens - lB;
    };
    double q_new = prop_density(zp, zc);
    double q_old = prop_density(zc, zp);
    double acc = lp_new - lp_old + q_old - q_new;
    for (int k = 0; k < K; k++) if (zc[k] <= 0) acc = 1;     // :170-174
    bool take = std::log(u[i]) < acc;
    if (acc_out) acc_out[i] = acc;
    if (accepted) accepted[i] = take ? 1 : 0;
    for (int k = 0; k < K; k++) Z_out[(size_t)k * n + i] = take ? zp[k] : zc[k];
  }
  return 0;
}

// ================================================================= chi sweep
// updateChi UpdateChi.h:19-64 (MV :138-174, cov-adj :242-290, MV cov-adj :370-411; tempered
// :116-119): for m in order, with chi(i,n) for n<m already updated.
extern "C" int orc_update_chi(const orc_data* d, const orc_state* s, double beta, const double* eps,
                              double* chi_out) {
  View v(d, s);
  PointEval pe(v);
  const int n = v.n, K = v.K, M = v.M;
  std::vector<double> cur(M);
  for (int i = 0; i < n; i++) {
    if (v.D) pe.set_function(i);
    for (int m = 0; m < M; m++) cur[m] = v.chi(i, m);
    const int64_t T = v.npts(i);
    for (int m = 0; m < M; m++) {
      double w = 0, W = 0;
      for (int64_t l = 0; l < T; l++) {
        pe.eval(i, l);
        double ph = 0;
        for (int k = 0; k < K; k++) ph += v.Z(i, k) * pe.fn[(size_t)k * M + m];
        w += ph * v.y(i, l);
        W += ph * ph;
        for (int k = 0; k < K; k++) {
          if (v.Z(i, k) != 0) {
            w -= v.Z(i, k) * ph * pe.an[k];
            for (int q = 0; q < M; q++)
              if (q != m) w -= v.Z(i, k) * ph * cur[q] * pe.fn[(size_t)k * M + q];
          }
        }
      }
      w = (w * beta) / s->sigma_sq;
      W = 1 + ((W * beta) / s->sigma_sq);
      W = 1 / W;
      cur[m] = W * w + std::sqrt(W) * eps[(size_t)m * n + i];   // R::rnorm(mu, sd) = mu + sd*N(0,1)
    }
    for (int m = 0; m < M; m++) chi_out[(size_t)m * n + i] = cur[m];
  }
  return 0;
}

// ================================================================= residual sum of squares
// the data pass shared by updateSigma (UpdateSigma.h:36-50) and calcLikelihood
// (CalculateLikelihood.h:28-42).
extern "C" int orc_ssr(const orc_data* d, const orc_state* s, double* ssr, double* sum_half,
                       double* n_points) {
  View v(d, s);
  PointEval pe(v);
  const int n = v.n, K = v.K, M = v.M;
  double acc = 0, half = 0, npts = 0;
  for (int i = 0; i < n; i++) {
    if (v.D) pe.set_function(i);
    const int64_t T = v.npts(i);
    for (int64_t l = 0; l < T; l++) {
      pe.eval(i, l);
      double b = v.y(i, l);
      for (int k = 0; k < K; k++) {
        if (v.Z(i, k) != 0) {
          b -= v.Z(i, k) * pe.an[k];
          for (int m = 0; m < M; m++) b -= v.Z(i, k) * v.chi(i, m) * pe.fn[(size_t)k * M + m];
        }
      }
      acc += b * b;
    }
    half += (double)(T / 2);          // integer division, UpdateSigma.h:49
    npts += (double)T;
  }
  if (v.ident) half = (double)(((int64_t)n * v.P) / 2);   // y_obs.n_elem / 2, UpdateSigma.h:150
  *ssr = acc;
  if (sum_half) *sum_half = half;
  if (n_points) *n_points = npts;
  return 0;
}

// sigma^2 | rest: 1 / rgamma(a, 1/b)  (UpdateSigma.h:51-53; tempered :98-107)
extern "C" int orc_update_sigma(const orc_data* d, const orc_state* s, double alpha0, double beta0,
                                double beta, int tempered, double gdraw, double* sigma_out,
                                double* shape_out, double* rate_out) {
  double ssr, half, npts;
  orc_ssr(d, s, &ssr, &half, &npts);
  double a, b1;
  if (tempered) { a = (beta * npts) / 2 + alpha0; b1 = (beta / 2) * ssr + beta0; }
  else          { a = half + alpha0;             b1 = 0.5 * ssr + beta0; }
  double scale = 1 / b1;
  double r = scale * gdraw;            // R::rgamma(a, scale) = scale * Gamma(a, 1)
  *sigma_out = 1 / r;
  if (shape_out) *shape_out = a;
  if (rate_out) *rate_out = b1;
  return 0;
}

// calcLikelihood: functional sums R::dnorm(y, mean, sqrt(sigma), log) per point
// (CalculateLikelihood.h:40); MV uses the closed form with floor(R/2) (:155).
extern "C" int orc_loglik(const orc_data* d, const orc_state* s, double* loglik) {
  View v(d, s);
  PointEval pe(v);
  const int n = v.n, K = v.K, M = v.M;
  const double sd = std::sqrt(s->sigma_sq);
  const double LN_SQRT_2PI = 0.918938533204672741780329736406;
  double ll = 0;
  for (int i = 0; i < n; i++) {
    if (v.D) pe.set_function(i);
    const int64_t T = v.npts(i);
    double ss = 0;
    for (int64_t l = 0; l < T; l++) {
      pe.eval(i, l);
      double mean = 0;
      for (int k = 0; k < K; k++) {
        if (v.Z(i, k) != 0) {
          mean += v.Z(i, k) * pe.an[k];
          for (int m = 0; m < M; m++) mean += v.Z(i, k) * v.chi(i, m) * pe.fn[(size_t)k * M + m];
        }
      }
      if (v.ident) { double r = v.y(i, l) - mean; ss += r * r; }
      else { double x = (v.y(i, l) - mean) / sd; ll += -(LN_SQRT_2PI + 0.5 * x * x + std::log(sd)); }
    }
    if (v.ident)
      ll = ll - ((double)(v.P / 2) * std::log(2 * M_PI * s->sigma_sq)) - ((1 / (s->sigma_sq * 2)) * ss);
  }
  *loglik = ll;
  return 0;
}

// Marginal (chi integrated out) log-likelihood of every function for ONE stored iteration: the summand of
// calcLikelihoodCPO (CalculateLikelihood.h:360-375):
//   mean_i = sum_k Z_ik B_i (nu_k + eta_k x_i)
//   cov_i  = sum_m (B_i u_im)(B_i u_im)' + sigma^2 I,  u_im = sum_k Z_ik (phi_km + xi_km x_i)
//            (:364-369 sums Z_ik Z_ik1 B (.)(.)' B' over k, k1: the same rank-one terms)
//   logl_i = -(n_i / 2) log(2 pi) - log det(cov_i) / 2 - (y_i - mean_i)' cov_i^{-1} (y_i - mean_i) / 2
// evaluated densely (n_i x n_i covariance, Cholesky), as the reference does.  The CPO itself is
//   CPO_i = log L + min_l logl_il - log sum_l exp(min_l logl_il - logl_il)          (:376-382)
// over the L retained iterations (the tests restate that line in numpy).
extern "C" int orc_marginal_loglik(const orc_data* d, const orc_state* s, double* logl_out) {
  View v(d, s);
  PointEval pe(v);
  const int n = v.n, K = v.K, M = v.M;
  for (int i = 0; i < n; i++) {
    if (v.D) pe.set_function(i);
    const int64_t T = v.npts(i);
    std::vector<double> res((size_t)T), U((size_t)T * std::max(M, 1)), cov((size_t)T * T, 0.0), L;
    for (int64_t l = 0; l < T; l++) {
      pe.eval(i, l);
      double mean = 0;
      for (int k = 0; k < K; k++) mean += v.Z(i, k) * pe.an[k];
      res[l] = v.y(i, l) - mean;
      for (int m = 0; m < M; m++) {
        double u = 0;
        for (int k = 0; k < K; k++) u += v.Z(i, k) * pe.fn[(size_t)k * M + m];
        U[(size_t)m * T + l] = u;
      }
    }
    for (int m = 0; m < M; m++)
      for (int64_t c = 0; c < T; c++)
        for (int64_t r = 0; r < T; r++) cov[(size_t)c * T + r] += U[(size_t)m * T + r] * U[(size_t)m * T + c];
    for (int64_t l = 0; l < T; l++) cov[(size_t)l * T + l] += s->sigma_sq;
    L.assign((size_t)T * T, 0.0);
    if (orc_chol_lower((int)T, cov.data(), L.data())) return 1;
    double logdet = 0;
    for (int64_t l = 0; l < T; l++) logdet += 2 * std::log(L[(size_t)l * T + l]);
    // w = L^{-1} res, quadratic form = |w|^2
    std::vector<double> w(res);
    double quad = 0;
    for (int64_t r = 0; r < T; r++) {
      double t = w[r];
      for (int64_t c = 0; c < r; c++) t -= L[(size_t)c * T + r] * w[c];
      w[r] = t / L[(size_t)r * T + r];
      quad += w[r] * w[r];
    }
    logl_out[i] = -(0.5 * (double)T) * std::log(2 * M_PI) - 0.5 * logdet - 0.5 * quad;
  }
  return 0;
}

// ================================================================= Gaussian block updates
static std::vector<double> scaled(const double* Pmat, int P, double c) {
  std::vector<double> pr((size_t)P * P);
  for (size_t e = 0; e < pr.size(); e++) pr[e] = c * Pmat[e];
  return pr;
}
static std::vector<double> ident_scaled(int P, double c) {
  std::vector<double> pr((size_t)P * P, 0.0);
  for (int p = 0; p < P; p++) pr[(size_t)p * P + p] = c;
  return pr;
}

// updateNu UpdateNu.h:24-74 (MV :160-204 prior (1/tau_j) I; cov-adj :287-344; MV cov-adj :443-493)
extern "C" int orc_update_nu(const orc_data* d, const orc_state* s, const double* tau,
                             const double* Pmat, double beta, const double* z, double* nu_out) {
  View v(d, s);
  BlockEngine be(v);
  for (int j = 0; j < v.K; j++) {
    std::vector<double> prior = v.ident ? ident_scaled(v.P, 1 / tau[j]) : scaled(Pmat, v.P, tau[j]);
    int rc = draw_block(be, {B_NU, j, 0, 0}, prior, true, beta, s->sigma_sq, z + (size_t)j * v.P);
    if (rc) return rc;
  }
  std::copy(v.nu.begin(), v.nu.end(), nu_out);
  return 0;
}

// updatePhi UpdatePhi.h:23-89 (MV :190-249; cov-adj :351-425; MV cov-adj :540-605)
extern "C" int orc_update_phi(const orc_data* d, const orc_state* s, const double* gamma,
                              const double* tilde_tau, double beta, const double* z,
                              double* Phi_out) {
  View v(d, s);
  BlockEngine be(v);
  const int K = v.K, P = v.P, M = v.M;
  int blk = 0;
  for (int j = 0; j < K; j++)
    for (int m = 0; m < M; m++, blk++) {
      std::vector<double> prior((size_t)P * P, 0.0);
      for (int p = 0; p < P; p++)
        prior[(size_t)p * P + p] = tilde_tau[(size_t)m * K + j] * gamma[((size_t)m * P + p) * K + j];
      int rc = draw_block(be, {B_PHI, j, m, 0}, prior, false, beta, s->sigma_sq, z + (size_t)blk * P);
      if (rc) return rc;
    }
  std::copy(v.Phi.begin(), v.Phi.end(), Phi_out);
  return 0;
}

// updateEta UpdateEta.h:28-94 (MV :203-262): d outer, j inner
extern "C" int orc_update_eta(const orc_data* d, const orc_state* s, const double* tau_eta,
                              const double* Pmat, double beta, const double* z, double* eta_out) {
  View v(d, s);
  BlockEngine be(v);
  int blk = 0;
  for (int dd = 0; dd < v.D; dd++)
    for (int j = 0; j < v.K; j++, blk++) {
      double te = tau_eta[(size_t)dd * v.K + j];
      std::vector<double> prior = v.ident ? ident_scaled(v.P, 1 / te) : scaled(Pmat, v.P, te);
      int rc = draw_block(be, {B_ETA, j, 0, dd}, prior, true, beta, s->sigma_sq, z + (size_t)blk * v.P);
      if (rc) return rc;
    }
  std::copy(v.eta.begin(), v.eta.end(), eta_out);
  return 0;
}

// updateXiCovariateAdj UpdateXi.h:26-93 (MV :201-260): order j, m, d
extern "C" int orc_update_xi(const orc_data* d, const orc_state* s, const double* gamma_xi,
                             const double* tilde_tau_xi, double beta, const double* z,
                             double* xi_out) {
  View v(d, s);
  BlockEngine be(v);
  const int K = v.K, P = v.P, M = v.M, D = v.D;
  int blk = 0;
  for (int j = 0; j < K; j++)
    for (int m = 0; m < M; m++)
      for (int dd = 0; dd < D; dd++, blk++) {
        std::vector<double> prior((size_t)P * P, 0.0);
        for (int p = 0; p < P; p++)
          prior[(size_t)p * P + p] = tilde_tau_xi[((size_t)dd * M + m) * K + j] *
                                     gamma_xi[(size_t)j * P * D * M + ((size_t)m * D + dd) * P + p];
        int rc = draw_block(be, {B_XI, j, m, dd}, prior, false, beta, s->sigma_sq, z + (size_t)blk * P);
        if (rc) return rc;
      }
  std::copy(v.xi.begin(), v.xi.end(), xi_out);
  return 0;
}

extern "C" int orc_pinv_sym(int n, const double* A, double* Ainv) { return pinv_sym(n, A, Ainv) ? 0 : 1; }
extern "C" int orc_inv(int n, const double* A, double* Ainv) { return inv_general(n, A, Ainv) ? 0 : 1; }
extern "C" int orc_chol_lower(int n, const double* A, double* L) { return chol_lower(n, A, L) ? 0 : 1; }
