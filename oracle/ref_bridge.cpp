// ref_bridge.cpp -- flat C entry points onto the REFERENCE's own update functions.
// TEST INFRASTRUCTURE ONLY.  The reference headers are compiled from where they lie
// (-I /root/reference/inst/include/BayesFMMM, see oracle/Makefile `make ref`) against the
// minimal Armadillo/Rmath stand-in in oracle/shim/.  Nothing from /root/reference is
// copied into this repository; this file only marshals flat arrays into the containers the
// reference functions take, calls them, and copies the results back.
//
// Signatures mirror bfmmm_oracle.h (ref_* instead of orc_*), so tests can run the
// restatement and the reference on the same inputs and the same injected draws
// (ref_tape_push).
#include <RcppArmadillo.h>
#include <truncnorm.h>

#include "Distributions.h"
#include "UpdateMixedMembership.h"
#include "UpdateChi.h"
#include "UpdateSigma.h"
#include "CalculateLikelihood.h"
#include "CalculateTTAcceptance.h"
#include "UpdateNu.h"
#include "UpdatePhi.h"
#include "UpdateEta.h"
#include "UpdateXi.h"
#include "UpdatePi.h"
#include "UpdateAlpha3.h"
#include "UpdateTau.h"
#include "UpdateDelta.h"
#include "UpdateGamma.h"
#include "UpdateA.h"

#include "bfmmm_oracle.h"

namespace {
using arma::uword;

struct Ctx {
  int n, K, P, M, D;
  bool ident;
  arma::field<arma::vec> y_f;
  arma::field<arma::mat> B_f;
  arma::mat y_mv, X, nu, Z, chi;
  arma::cube Phi, eta;
  arma::field<arma::cube> xi;     // rows x K, every row identical (see SURVEY finding 0.6)
  double sigma;

  Ctx(const orc_data* d, const orc_state* s, int xi_rows = 1) {
    n = d->n; K = d->K; P = d->P; M = d->M; D = d->D; ident = d->identity_basis != 0;
    if (ident) {
      y_mv = arma::mat(n, P);
      std::copy(d->y, d->y + (size_t)n * P, y_mv.memptr());
    } else {
      y_f = arma::field<arma::vec>(n, 1);
      B_f = arma::field<arma::mat>(n, 1);
      for (int i = 0; i < n; i++) {
        int64_t T = d->off[i + 1] - d->off[i];
        y_f(i, 0) = arma::vec(T);
        B_f(i, 0) = arma::mat(T, P);
        for (int64_t l = 0; l < T; l++) {
          y_f(i, 0)(l) = d->y[d->off[i] + l];
          for (int p = 0; p < P; p++) B_f(i, 0)(l, p) = d->B[(size_t)(d->off[i] + l) * P + p];
        }
      }
    }
    nu = arma::mat(K, P); std::copy(s->nu, s->nu + (size_t)K * P, nu.memptr());
    Phi = arma::cube(K, P, M);
    for (int m = 0; m < M; m++) std::copy(s->Phi + (size_t)m * K * P, s->Phi + (size_t)(m + 1) * K * P, Phi.slice(m).memptr());
    Z = arma::mat(n, K); std::copy(s->Z, s->Z + (size_t)n * K, Z.memptr());
    chi = arma::mat(n, M); if (M) std::copy(s->chi, s->chi + (size_t)n * M, chi.memptr());
    sigma = s->sigma_sq;
    if (D) {
      X = arma::mat(n, D); std::copy(d->X, d->X + (size_t)n * D, X.memptr());
      eta = arma::cube(P, D, K);
      for (int k = 0; k < K; k++) std::copy(s->eta + (size_t)k * P * D, s->eta + (size_t)(k + 1) * P * D, eta.slice(k).memptr());
      xi = arma::field<arma::cube>(xi_rows, K);
      for (int r = 0; r < xi_rows; r++)
        for (int k = 0; k < K; k++) {
          xi(r, k) = arma::cube(P, D, M);
          for (int m = 0; m < M; m++)
            std::copy(s->xi + (size_t)k * P * D * M + (size_t)m * P * D, s->xi + (size_t)k * P * D * M + (size_t)(m + 1) * P * D,
                      xi(r, k).slice(m).memptr());
        }
    }
  }
};

void cube_out(const arma::cube& c, double* out) {
  size_t sl = (size_t)c.n_rows * c.n_cols;
  for (uword s = 0; s < c.n_slices; s++) std::copy(c.slice(s).memptr(), c.slice(s).memptr() + sl, out + s * sl);
}
arma::cube cube_in(const double* p, int r, int c, int s) {
  arma::cube q(r, c, s);
  for (int i = 0; i < s; i++) std::copy(p + (size_t)i * r * c, p + (size_t)(i + 1) * r * c, q.slice(i).memptr());
  return q;
}
arma::mat mat_in(const double* p, int r, int c) { arma::mat m(r, c); std::copy(p, p + (size_t)r * c, m.memptr()); return m; }
arma::vec vec_in(const double* p, int n) { arma::vec v(n); std::copy(p, p + n, v.memptr()); return v; }
}  // namespace

extern "C" {

void ref_tape_clear() { shim::tape().q.clear(); shim::tape().popped = 0; }
void ref_tape_push(const double* v, int64_t n) { for (int64_t i = 0; i < n; i++) shim::tape().q.push_back(v[i]); }
int64_t ref_tape_left() { return (int64_t)shim::tape().q.size(); }
void ref_seed(uint64_t s) { shim::tape().rng.seed(s); }

// Draw order per function i: K gammas (rdirichlet) then 1 uniform -- the tape must be laid out so.
int ref_update_z(const orc_data* d, const orc_state* s, const double* pi, double alpha3,
                 double a_Z_PM, double beta, int tempered, double* Z_out) {
  try {
    Ctx c(d, s, d->D ? d->n : 1);
    arma::cube Z(c.n, c.K, 1); Z.slice(0) = c.Z;
    arma::vec piv = vec_in(pi, c.K), Z_ph = arma::zeros(c.K);
    using namespace BayesFMMM;
    if (!c.ident && !c.D) {
      if (tempered) updateZTempered_PM(beta, c.y_f, c.B_f, c.Phi, c.nu, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, Z_ph, Z);
      else updateZ_PM(c.y_f, c.B_f, c.Phi, c.nu, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, Z_ph, Z);
    } else if (c.ident && !c.D) {
      if (tempered) updateZTempered_MMMV(beta, c.y_mv, c.Phi, c.nu, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, Z_ph, Z);
      else updateZ_MMMV(c.y_mv, c.Phi, c.nu, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, Z_ph, Z);
    } else if (!c.ident) {
      if (tempered) updateZTempered_PMCovariateAdj(beta, c.y_f, c.B_f, c.Phi, c.xi, c.nu, c.eta, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, c.X, Z_ph, Z);
      else updateZ_PMCovariateAdj(c.y_f, c.B_f, c.Phi, c.xi, c.nu, c.eta, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, c.X, Z_ph, Z);
    } else {
      if (tempered) updateZTempered_MMMVCovariateAdj(beta, c.y_mv, c.Phi, c.xi, c.nu, c.eta, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, c.X, Z_ph, Z);
      else updateZ_MMMVCovariateAdj(c.y_mv, c.Phi, c.xi, c.nu, c.eta, c.chi, piv, c.sigma, 0, 1, alpha3, a_Z_PM, c.X, Z_ph, Z);
    }
    std::copy(Z.slice(0).memptr(), Z.slice(0).memptr() + (size_t)c.n * c.K, Z_out);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_z: " << e.what() << "\n"; return 1; }
}

// Draw order: for i, for m: one normal.
int ref_update_chi(const orc_data* d, const orc_state* s, double beta, int tempered, double* chi_out) {
  try {
    Ctx c(d, s);
    arma::cube chi(c.n, c.M, 1); chi.slice(0) = c.chi;
    using namespace BayesFMMM;
    if (!c.ident && !c.D) {
      if (tempered) updateChiTempered(beta, c.y_f, c.B_f, c.Phi, c.nu, c.Z, c.sigma, 0, 1, chi);
      else updateChi(c.y_f, c.B_f, c.Phi, c.nu, c.Z, c.sigma, 0, 1, chi);
    } else if (c.ident && !c.D) {
      if (tempered) updateChiTemperedMV(beta, c.y_mv, c.Phi, c.nu, c.Z, c.sigma, 0, 1, chi);
      else updateChiMV(c.y_mv, c.Phi, c.nu, c.Z, c.sigma, 0, 1, chi);
    } else if (!c.ident) {
      if (tempered) updateChiTemperedCovariateAdj(beta, c.y_f, c.B_f, c.Phi, c.xi, c.nu, c.eta, c.Z, c.sigma, 0, 1, c.X, chi);
      else updateChiCovariateAdj(c.y_f, c.B_f, c.Phi, c.xi, c.nu, c.eta, c.Z, c.sigma, 0, 1, c.X, chi);
    } else {
      if (tempered) updateChiTemperedMVCovariateAdj(beta, c.y_mv, c.Phi, c.xi, c.nu, c.eta, c.Z, c.sigma, 0, 1, c.X, chi);
      else updateChiMVCovariateAdj(c.y_mv, c.Phi, c.xi, c.nu, c.eta, c.Z, c.sigma, 0, 1, c.X, chi);
    }
    std::copy(chi.slice(0).memptr(), chi.slice(0).memptr() + (size_t)c.n * c.M, chi_out);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_chi: " << e.what() << "\n"; return 1; }
}

// Draw order: one standard gamma.
int ref_update_sigma(const orc_data* d, const orc_state* s, double alpha0, double beta0, double beta,
                     int tempered, double* sigma_out) {
  try {
    Ctx c(d, s);
    arma::vec sg = arma::ones(1);
    using namespace BayesFMMM;
    if (!c.ident && !c.D) {
      if (tempered) updateSigmaTempered(beta, c.y_f, c.B_f, alpha0, beta0, c.nu, c.Phi, c.Z, c.chi, 0, 1, sg);
      else updateSigma(c.y_f, c.B_f, alpha0, beta0, c.nu, c.Phi, c.Z, c.chi, 0, 1, sg);
    } else if (c.ident && !c.D) {
      if (tempered) updateSigmaTemperedMV(beta, c.y_mv, alpha0, beta0, c.nu, c.Phi, c.Z, c.chi, 0, 1, sg);
      else updateSigmaMV(c.y_mv, alpha0, beta0, c.nu, c.Phi, c.Z, c.chi, 0, 1, sg);
    } else if (!c.ident) {
      if (tempered) updateSigmaTemperedCovariateAdj(beta, c.y_f, c.B_f, alpha0, beta0, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, 1, c.X, sg);
      else updateSigmaCovariateAdj(c.y_f, c.B_f, alpha0, beta0, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, 1, c.X, sg);
    } else {
      if (tempered) updateSigmaTemperedMVCovariateAdj(beta, c.y_mv, alpha0, beta0, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, 1, c.X, sg);
      else updateSigmaMVCovariateAdj(c.y_mv, alpha0, beta0, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, 1, c.X, sg);
    }
    *sigma_out = sg(0);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_sigma: " << e.what() << "\n"; return 1; }
}

int ref_loglik(const orc_data* d, const orc_state* s, double* ll) {
  try {
    Ctx c(d, s);
    using namespace BayesFMMM;
    if (!c.ident && !c.D) *ll = calcLikelihood(c.y_f, c.B_f, c.nu, c.Phi, c.Z, c.chi, c.sigma);
    else if (c.ident && !c.D) *ll = calcLikelihoodMV(c.y_mv, c.nu, c.Phi, c.Z, c.chi, c.sigma);
    else if (!c.ident) *ll = calcLikelihoodCovariateAdj(c.y_f, c.B_f, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, c.X, c.sigma);
    else *ll = calcLikelihoodMVCovariateAdj(c.y_mv, c.nu, c.eta, c.Phi, c.xi, c.Z, c.chi, 0, c.X, c.sigma);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_loglik: " << e.what() << "\n"; return 1; }
}

// calcLikelihoodCPO (CalculateLikelihood.h:344-385) over `iters` stored iterations (states[l]); with one
// iteration and burnin 0 the returned CPO_i is logl_i itself.  Without covariates the caller of the
// reference passes a zero covariate column and zero eta / xi (src/PostProcessing.cpp:6455-6475).
int ref_cpo(const orc_data* d, const orc_state* states, int iters, double burnin_prop, double* cpo_out) {
  try {
    if (d->identity_basis) return 1;
    const int n = d->n, K = d->K, P = d->P, M = d->M, D = d->D, De = D ? D : 1;
    Ctx c0(d, &states[0]);
    arma::cube nu(K, P, iters), Z(n, K, iters), chi(n, M, iters);
    arma::field<arma::cube> eta(iters, 1), Phi(iters, 1), xi(iters, K);
    arma::vec sigma(iters);
    for (int l = 0; l < iters; l++) {
      Ctx c(d, &states[l]);
      nu.slice(l) = c.nu; Z.slice(l) = c.Z; chi.slice(l) = c.chi; sigma(l) = c.sigma; Phi(l, 0) = c.Phi;
      if (D) { eta(l, 0) = c.eta; for (int k = 0; k < K; k++) xi(l, k) = c.xi(0, k); }
      else { eta(l, 0) = arma::zeros(P, 1, K); for (int k = 0; k < K; k++) xi(l, k) = arma::zeros(P, 1, M); }
    }
    arma::mat X = D ? c0.X : arma::mat(arma::zeros(n, De));
    arma::vec out = BayesFMMM::calcLikelihoodCPO(c0.y_f, c0.B_f, nu, eta, Phi, xi, Z, chi, X, sigma, iters, burnin_prop);
    for (int i = 0; i < n; i++) cpo_out[i] = out(i);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_cpo: " << e.what() << "\n"; return 1; }
}

// Draw order: for j: P normals.
int ref_update_nu(const orc_data* d, const orc_state* s, const double* tau, const double* Pmat,
                  double beta, int tempered, double* nu_out) {
  try {
    Ctx c(d, s);
    arma::cube nu(c.K, c.P, 1); nu.slice(0) = c.nu;
    arma::vec tauv = vec_in(tau, c.K), b_1 = arma::zeros(c.P);
    arma::mat B_1 = arma::zeros(c.P, c.P), Pm;
    if (!c.ident) Pm = mat_in(Pmat, c.P, c.P);
    using namespace BayesFMMM;
    if (!c.ident && !c.D) {
      if (tempered) updateNuTempered(beta, c.y_f, c.B_f, tauv, c.Phi, c.Z, c.chi, c.sigma, 0, 1, Pm, b_1, B_1, nu);
      else updateNu(c.y_f, c.B_f, tauv, c.Phi, c.Z, c.chi, c.sigma, 0, 1, Pm, b_1, B_1, nu);
    } else if (c.ident && !c.D) {
      if (tempered) updateNuTemperedMV(beta, c.y_mv, tauv, c.Phi, c.Z, c.chi, c.sigma, 0, 1, b_1, B_1, nu);
      else updateNuMV(c.y_mv, tauv, c.Phi, c.Z, c.chi, c.sigma, 0, 1, b_1, B_1, nu);
    } else if (!c.ident) {
      if (tempered) updateNuTemperedCovariateAdj(beta, c.y_f, c.B_f, tauv, c.Phi, c.xi, c.eta, c.Z, c.chi, c.sigma, 0, 1, Pm, c.X, b_1, B_1, nu);
      else updateNuCovariateAdj(c.y_f, c.B_f, tauv, c.Phi, c.xi, c.eta, c.Z, c.chi, c.sigma, 0, 1, Pm, c.X, b_1, B_1, nu);
    } else {
      if (tempered) updateNuTemperedMVCovariateAdj(beta, c.y_mv, tauv, c.Phi, c.xi, c.eta, c.Z, c.chi, c.sigma, 0, 1, c.X, b_1, B_1, nu);
      else updateNuMVCovariateAdj(c.y_mv, tauv, c.Phi, c.xi, c.eta, c.Z, c.chi, c.sigma, 0, 1, c.X, b_1, B_1, nu);
    }
    std::copy(nu.slice(0).memptr(), nu.slice(0).memptr() + (size_t)c.K * c.P, nu_out);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_nu: " << e.what() << "\n"; return 1; }
}

// Draw order: for j, for m: P normals.
int ref_update_phi(const orc_data* d, const orc_state* s, const double* gamma, const double* tilde_tau,
                   double beta, int tempered, double* Phi_out) {
  try {
    Ctx c(d, s);
    arma::field<arma::cube> Phi(1, 1); Phi(0, 0) = c.Phi;
    arma::cube gam = cube_in(gamma, c.K, c.P, c.M);
    arma::mat tt = mat_in(tilde_tau, c.K, c.M), M_1 = arma::zeros(c.P, c.P);
    arma::vec m_1 = arma::zeros(c.P);
    using namespace BayesFMMM;
    if (!c.ident && !c.D) {
      if (tempered) updatePhiTempered(beta, c.y_f, c.B_f, c.nu, gam, tt, c.Z, c.chi, c.sigma, 0, 1, m_1, M_1, Phi);
      else updatePhi(c.y_f, c.B_f, c.nu, gam, tt, c.Z, c.chi, c.sigma, 0, 1, m_1, M_1, Phi);
    } else if (c.ident && !c.D) {
      if (tempered) updatePhiTemperedMV(beta, c.y_mv, c.nu, gam, tt, c.Z, c.chi, c.sigma, 0, 1, m_1, M_1, Phi);
      else updatePhiMV(c.y_mv, c.nu, gam, tt, c.Z, c.chi, c.sigma, 0, 1, m_1, M_1, Phi);
    } else if (!c.ident) {
      if (tempered) updatePhiTemperedCovariateAdj(beta, c.y_f, c.B_f, c.nu, c.eta, gam, tt, c.xi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, Phi);
      else updatePhiCovariateAdj(c.y_f, c.B_f, c.nu, c.eta, gam, tt, c.xi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, Phi);
    } else {
      if (tempered) updatePhiTemperedMVCovariateAdj(beta, c.y_mv, c.nu, c.eta, gam, tt, c.xi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, Phi);
      else updatePhiMVCovariateAdj(c.y_mv, c.nu, c.eta, gam, tt, c.xi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, Phi);
    }
    cube_out(Phi(0, 0), Phi_out);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_phi: " << e.what() << "\n"; return 1; }
}

// Draw order: for d, for j: P normals.
int ref_update_eta(const orc_data* d, const orc_state* s, const double* tau_eta, const double* Pmat,
                   double beta, int tempered, double* eta_out) {
  try {
    Ctx c(d, s);
    arma::field<arma::cube> eta(1, 1); eta(0, 0) = c.eta;
    arma::mat te = mat_in(tau_eta, c.K, c.D), B_1 = arma::zeros(c.P, c.P), Pm;
    arma::vec b_1 = arma::zeros(c.P);
    if (!c.ident) Pm = mat_in(Pmat, c.P, c.P);
    using namespace BayesFMMM;
    if (!c.ident) {
      if (tempered) updateEtaTempered(beta, c.y_f, c.B_f, te, c.Phi, c.xi, c.nu, c.Z, c.chi, c.sigma, 0, 1, Pm, c.X, b_1, B_1, eta);
      else updateEta(c.y_f, c.B_f, te, c.Phi, c.xi, c.nu, c.Z, c.chi, c.sigma, 0, 1, Pm, c.X, b_1, B_1, eta);
    } else {
      if (tempered) updateEtaTemperedMV(beta, c.y_mv, te, c.Phi, c.xi, c.nu, c.Z, c.chi, c.sigma, 0, 1, c.X, b_1, B_1, eta);
      else updateEtaMV(c.y_mv, te, c.Phi, c.xi, c.nu, c.Z, c.chi, c.sigma, 0, 1, c.X, b_1, B_1, eta);
    }
    cube_out(eta(0, 0), eta_out);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_eta: " << e.what() << "\n"; return 1; }
}

// Draw order: for j, m, d: P normals.  gamma_xi: K cubes P x D x M; tilde_tau_xi: K x M x D.
int ref_update_xi(const orc_data* d, const orc_state* s, const double* gamma_xi, const double* tilde_tau_xi,
                  double beta, int tempered, double* xi_out) {
  try {
    Ctx c(d, s);
    arma::field<arma::cube> gx(1, c.K);
    for (int k = 0; k < c.K; k++) gx(0, k) = cube_in(gamma_xi + (size_t)k * c.P * c.D * c.M, c.P, c.D, c.M);
    arma::cube tt = cube_in(tilde_tau_xi, c.K, c.M, c.D);
    arma::mat M_1 = arma::zeros(c.P, c.P); arma::vec m_1 = arma::zeros(c.P);
    using namespace BayesFMMM;
    if (!c.ident) {
      if (tempered) updateXiTemperedCovariateAdj(beta, c.y_f, c.B_f, c.nu, c.eta, gx, tt, c.Phi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, c.xi);
      else updateXiCovariateAdj(c.y_f, c.B_f, c.nu, c.eta, gx, tt, c.Phi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, c.xi);
    } else {
      if (tempered) updateXiTemperedMVCovariateAdj(beta, c.y_mv, c.nu, c.eta, gx, tt, c.Phi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, c.xi);
      else updateXiMVCovariateAdj(c.y_mv, c.nu, c.eta, gx, tt, c.Phi, c.Z, c.chi, c.sigma, c.X, 0, 1, m_1, M_1, c.xi);
    }
    for (int k = 0; k < c.K; k++) cube_out(c.xi(0, k), xi_out + (size_t)k * c.P * c.D * c.M);
    return 0;
  } catch (std::exception& e) { std::cerr << "ref_update_xi: " << e.what() << "\n"; return 1; }
}

// ---- host-side prior updates (small; used to check the product's host loop) ----
// updatePi_PM: K gammas + 1 uniform
int ref_update_pi(int n, int K, const double* Z, const double* c_hyp, double alpha3, double a_pi_PM,
                  const double* pi_in, double* pi_out) {
  arma::mat Zm = mat_in(Z, n, K), pi(K, 1);
  std::copy(pi_in, pi_in + K, pi.memptr());
  arma::vec cv = vec_in(c_hyp, K), ph = arma::zeros(K);
  BayesFMMM::updatePi_PM(alpha3, Zm, cv, 0, 1, a_pi_PM, ph, pi);
  std::copy(pi.memptr(), pi.memptr() + K, pi_out);
  return 0;
}
// updateAlpha3: 1 uniform (truncnorm proposal) + 1 uniform
int ref_update_alpha3(int n, int K, const double* Z, const double* pi, double b, double var_alpha3,
                      double alpha3_in, double* alpha3_out) {
  arma::mat Zm = mat_in(Z, n, K);
  arma::vec piv = vec_in(pi, K), a3 = arma::ones(1);
  a3(0) = alpha3_in;
  BayesFMMM::updateAlpha3(piv, b, Zm, 0, 1, var_alpha3, a3);
  *alpha3_out = a3(0);
  return 0;
}
// updateTau / updateTauMV: K gammas
int ref_update_tau(int K, int P, const double* nu, const double* Pmat, double alpha, double beta, int mv,
                   double* tau_out) {
  arma::mat num = mat_in(nu, K, P), tau(1, K, arma::fill::ones);
  if (mv) BayesFMMM::updateTauMV(alpha, beta, num, 0, 1, tau);
  else { arma::mat Pm = mat_in(Pmat, P, P); BayesFMMM::updateTau(alpha, beta, num, 0, 1, Pm, tau); }
  std::copy(tau.memptr(), tau.memptr() + K, tau_out);
  return 0;
}
// updateTauEta / MV: tau_eta cube K x D x iters
int ref_update_tau_eta(int K, int P, int D, const double* eta, const double* Pmat, double alpha, double beta,
                       int mv, double* out) {
  arma::cube e = cube_in(eta, P, D, K), te(K, D, 1, arma::fill::ones);
  if (mv) BayesFMMM::updateTauEtaMV(alpha, beta, e, 0, 1, te);
  else { arma::mat Pm = mat_in(Pmat, P, P); BayesFMMM::updateTauEta(alpha, beta, e, 0, 1, Pm, te); }
  std::copy(te.slice(0).memptr(), te.slice(0).memptr() + (size_t)K * D, out);
  return 0;
}
// updateDelta: K*M gammas, uses delta in place
int ref_update_delta(int K, int P, int M, const double* Phi, const double* gamma, const double* A,
                     const double* delta_in, double* delta_out) {
  arma::cube phi = cube_in(Phi, K, P, M), gam = cube_in(gamma, K, P, M), delta(K, M, 1);
  std::copy(delta_in, delta_in + (size_t)K * M, delta.slice(0).memptr());
  arma::mat a = mat_in(A, K, 2);
  BayesFMMM::updateDelta(phi, gam, a, 0, 1, delta);
  std::copy(delta.slice(0).memptr(), delta.slice(0).memptr() + (size_t)K * M, delta_out);
  return 0;
}
// updateGamma: K*P*M gammas (order i, l, j)
int ref_update_gamma(int K, int P, int M, double nu_gamma, const double* delta, const double* Phi, double* gamma_out) {
  arma::cube phi = cube_in(Phi, K, P, M);
  arma::mat del = mat_in(delta, K, M);
  arma::field<arma::cube> gam(1, 1); gam(0, 0) = arma::cube(K, P, M, arma::fill::ones);
  BayesFMMM::updateGamma(nu_gamma, del, phi, 0, 1, gam);
  cube_out(gam(0, 0), gamma_out);
  return 0;
}
// updateA: per (j, i): 1 uniform (truncnorm proposal) + 1 uniform
int ref_update_A(int K, int M, double a1l, double b1l, double a2l, double b2l, const double* delta,
                 double ve1, double ve2, const double* A_in, double* A_out) {
  arma::mat del = mat_in(delta, K, M);
  arma::cube a(K, 2, 1);
  std::copy(A_in, A_in + (size_t)K * 2, a.slice(0).memptr());
  BayesFMMM::updateA(a1l, b1l, a2l, b2l, del, ve1, ve2, 0, 1, a);
  std::copy(a.slice(0).memptr(), a.slice(0).memptr() + (size_t)K * 2, A_out);
  return 0;
}

// updateDeltaXi: draws in order d, k, i.  xi/gamma_xi: K cubes P x D x M; A_xi K x 2 x D; delta K x M x D
int ref_update_delta_xi(int K, int P, int M, int D, const double* xi, const double* gamma_xi, const double* A_xi,
                        const double* delta_in, double* delta_out) {
  arma::field<arma::cube> xif(1, K), gx(1, K), del(1, 1);
  for (int k = 0; k < K; k++) {
    xif(0, k) = cube_in(xi + (size_t)k * P * D * M, P, D, M);
    gx(0, k) = cube_in(gamma_xi + (size_t)k * P * D * M, P, D, M);
  }
  del(0, 0) = cube_in(delta_in, K, M, D);
  arma::cube ax = cube_in(A_xi, K, 2, D);
  BayesFMMM::updateDeltaXi(xif, gx, ax, 0, 1, del);
  cube_out(del(0, 0), delta_out);
  return 0;
}
// updateGammaXi: draws in order k, d, p, m
int ref_update_gamma_xi(int K, int P, int M, int D, double nu_gamma, const double* delta_xi, const double* xi,
                        double* gamma_out) {
  arma::field<arma::cube> xif(1, K), gx(1, K);
  for (int k = 0; k < K; k++) {
    xif(0, k) = cube_in(xi + (size_t)k * P * D * M, P, D, M);
    gx(0, k) = arma::cube(P, D, M, arma::fill::ones);
  }
  arma::cube del = cube_in(delta_xi, K, M, D);
  BayesFMMM::updateGammaXi(nu_gamma, del, xif, 0, 1, gx);
  for (int k = 0; k < K; k++) cube_out(gx(0, k), gamma_out + (size_t)k * P * D * M);
  return 0;
}
// updateAXi: per (j, i, d): 1 uniform (proposal) + 1 uniform
int ref_update_A_xi(int K, int M, int D, double a1l, double b1l, double a2l, double b2l, const double* delta_xi,
                    double ve1, double ve2, const double* A_in, double* A_out) {
  arma::cube del = cube_in(delta_xi, K, M, D);
  arma::field<arma::cube> a(1, 1);
  a(0, 0) = cube_in(A_in, K, 2, D);
  BayesFMMM::updateAXi(a1l, b1l, a2l, b2l, del, ve1, ve2, 0, 1, a);
  cube_out(a(0, 0), A_out);
  return 0;
}

}  // extern "C"
