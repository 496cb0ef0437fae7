/*
 * bfmmm.h -- C ABI of the B200 (sm_100a) engine for BayesFMMM's per-iteration sampler.
 *
 * Drop-in seam: the calls the reference's host loops make into the observation-level updates
 * (SURVEY.md section 8b), e.g. in BFMMM_Theta inst/include/BayesFMMM/BFMMM.h:1261-1292 and
 * BFMMM_MTT_warm_start BFMMM.h:1500-1554:
 *     updateZ_PM, updatePhi, updateNu, updateSigma, updateChi, calcLikelihood (+ MV / covariate-
 *     adjusted / tempered twins).
 * A patched driver replaces each of those calls by the entry point named next to it below; all
 * small prior updates (tau, delta, gamma, A, pi, alpha_3) and the P x P solves stay in the
 * caller's Armadillo code.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is HOST memory unless the name ends in _dev; FP64; Armadillo (column-major)
 *     layouts, so `.memptr()` of the reference's containers can be passed directly:
 *       nu  K x P | Phi K x P x M | eta P x D x K | xi: K cubes P x D x M back to back
 *       Z n x K | chi n x M | X n x D
 *   - return value 0 = ok; anything else is an error, message via bfmmm_last_error().
 *     (The reference reports errors as C++ exceptions -> Rcpp::stop; the stub rethrows.)
 *   - the engine never keeps host pointers after a call returns.  Observations (y, grids, X)
 *     are copied to the device once at create time; Z and chi live on the device between calls.
 *   - calls on one handle must come from one thread at a time (R is single-threaded).
 *   - `beta` is the tempering temperature of the *Tempered twins (1.0 = untempered functions).
 *   - there is no CPU fallback: if no CUDA device / the kernels are unavailable, create fails.
 */
#ifndef BFMMM_H
#define BFMMM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bfmmm_engine bfmmm_engine;

enum {
  BFMMM_FUNCTIONAL = 0,  /* B_i from a B-spline / tensor basis; BFMMM_* and BHDFMMM_* drivers   */
  BFMMM_MULTIVARIATE = 1 /* B_i = I_P (P = data dimension); BMVMMM_* drivers                    */
};

typedef struct {
  int32_t model;        /* BFMMM_FUNCTIONAL | BFMMM_MULTIVARIATE                                   */
  int32_t n;            /* functions / observations held by THIS engine (this rank's shard)        */
  int32_t K, P, M, D;   /* features, basis dimension, eigenfunctions, covariates (0 = none)        */
  int32_t device;       /* CUDA device ordinal                                                     */
  int32_t common_grid;  /* functional: 1 = every function observed on the same T points            */
  int64_t T;            /* common grid: points per function; ragged: ignored                       */
  const int64_t* off;   /* ragged: n+1 offsets into y / B_rows; common grid: may be NULL           */
  const double* y;      /* functional: concatenated y_i (sum n_i); MV: n x P column-major          */
  const double* B;      /* functional common grid: T x P basis, ROW-major (row = grid point);      */
                        /* functional ragged: (sum n_i) x P row-major; MV: NULL.                   */
                        /* NULL for functional => evaluate the B-spline basis on the device from    */
                        /* the spline description below (replaces splines2::BSpline(...).basis())   */
  const double* t;      /* grid points: common T, or concatenated (sum n_i); only if B == NULL      */
  int32_t degree;       /* spline degree (basis_degree)                                            */
  int32_t n_internal;   /* number of internal knots                                                */
  const double* internal_knots;
  double boundary[2];
  const double* X;      /* n x D column-major covariates or NULL                                   */
  int64_t global_offset;/* index of this shard's first function in the whole data set (RNG ctr)    */
} bfmmm_config;

/* ---- lifetime ------------------------------------------------------------------------------ */
int bfmmm_create(const bfmmm_config* cfg, bfmmm_engine** out);
void bfmmm_destroy(bfmmm_engine* e);
const char* bfmmm_last_error(void);
/* number of kernels this library has launched since load (all engines) */
int64_t bfmmm_launch_count(void);
/* basis matrix the engine uses (functional): common grid T x P row-major. Replaces the
 * B_obs the drivers return in their result list (UserFunctions.cpp:887). */
int bfmmm_get_basis(bfmmm_engine* e, double* B_out);

/* ---- per-observation state (replaces the Z / chi chain slices the updates read and write) --- */
int bfmmm_set_state(bfmmm_engine* e, const double* Z, const double* chi);
int bfmmm_get_state(bfmmm_engine* e, double* Z, double* chi);     /* either may be NULL */
/* Overlapped read-back for drivers that keep every iteration's Z / chi slice (the reference's chain cubes,
 * BFMMM.h:1253-1298 writes slice i+1 every iteration).  _begin snapshots the current (Z, chi) device-side,
 * ordered after everything queued so far, and starts the device-to-host transfer on a separate copy
 * stream into the caller's buffers (page-locked memory for the transfer to be asynchronous), so the
 * next iteration's kernels run while the slice travels; it returns immediately.  _wait blocks until the
 * last transfer started has landed.  A new _begin waits (on the device) for the previous transfer. */
int bfmmm_get_state_begin(bfmmm_engine* e, double* Z, double* chi);   /* either may be NULL */
int bfmmm_get_state_wait(bfmmm_engine* e);
/* rows [i0, i0+count) only: Z count x K, chi count x M (column-major, leading dimension count) */
int bfmmm_get_state_rows(bfmmm_engine* e, int64_t i0, int64_t count, double* Z, double* chi);

/* sigma^2 drawn on the device right behind bfmmm_ssr_async (and the caller's all-reduce of the SSR slot):
 * sigma^2 = 1 / ((1/b1) Gamma(a)), b1 = scale_ssr * SSR + beta_0 (UpdateSigma.h:47-53, tempered :98-107), from
 * the counter-based stream (key, iteration, purpose).  The next bfmmm_update_chi_async reads it from device
 * memory, so no host round trip separates the SSR pass from the chi pass; bfmmm_sigma_wait returns
 * (SSR, sigma^2) as soon as the draw has happened (it does not wait for kernels queued behind it). */
int bfmmm_sigma_draw_async(bfmmm_engine* e, double a, double scale_ssr, double beta0, uint64_t key,
                           uint64_t iteration, uint32_t purpose);
int bfmmm_sigma_wait(bfmmm_engine* e, double* ssr, double* sigma_sq);

/* ---- post-processing on the device ------------------------------------------------------
 * calcLikelihoodCPO (CalculateLikelihood.h:344-385, called by ConditionalPredictiveOrdinates,
 * src/PostProcessing.cpp:6509): per stored iteration the marginal log-likelihood of every function with chi
 * integrated out, then CPO_i = log L + min_l logl_il - log sum_l exp(min_l logl_il - logl_il).  The caller
 * replays the stored iterations (bfmmm_set_state + bfmmm_set_globals per retained iteration, burn-in skipped)
 * and calls bfmmm_cpo_accumulate for each; bfmmm_cpo_get returns the CPO (log scale unless log_scale == 0). */
int bfmmm_marginal_loglik(bfmmm_engine* e, double* logl /* n */);
int bfmmm_cpo_reset(bfmmm_engine* e);
int bfmmm_cpo_accumulate(bfmmm_engine* e);
int bfmmm_cpo_get(bfmmm_engine* e, double* cpo /* n */, int log_scale);

/* device-side snapshot / restore of (Z, chi): a rejected tempered transition keeps the
 * pre-transition slice (BFMMM.h:1631-1651) */
int bfmmm_state_snapshot(bfmmm_engine* e);
int bfmmm_state_restore(bfmmm_engine* e);

/* ---- global parameters: pushed before the phases that read them ----------------------------- */
/* eta/xi may be NULL when D == 0 (M >= 1 always: bfmmm_create rejects M < 1). */
int bfmmm_set_globals(bfmmm_engine* e, const double* nu, const double* Phi, const double* eta,
                      const double* xi, double sigma_sq);

/* ---- the hot path --------------------------------------------------------------------------- */
/* updateZ_PM / updateZ_MMMV / *CovariateAdj / *Tempered (UpdateMixedMembership.h:131-185 ...).
 * Injected-draw mode (parity): gam = n x K raw Gamma(a_Z_PM*Z_ik, 1) draws, u = n uniforms.
 * Device-RNG mode: gam == NULL and u == NULL; draws come from Philox keyed by bfmmm_seed and the
 * global function index.  Outputs (any may be NULL): sum_log_Z[K] = sum_i log Z_ik of the NEW
 * state (feeds updatePi_PM / updateAlpha3, UpdatePi.h:45-49), n_accept. */
int bfmmm_update_z(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta,
                   const double* gam, const double* u, double* sum_log_Z, int64_t* n_accept);
/* updateChi and twins (UpdateChi.h:19-64 ...): sequential-in-m sweep.  eps = n x M N(0,1) draws
 * or NULL for device RNG.  ssr_after (may be NULL) = residual sum of squares with the NEW chi,
 * which is what the calcLikelihood call that follows in the driver loop needs (BFMMM.h:1291). */
int bfmmm_update_chi(bfmmm_engine* e, double beta, const double* eps, double* ssr_after);
/* the data pass of updateSigma / calcLikelihood (UpdateSigma.h:36-50, CalculateLikelihood.h:28-42):
 * ssr = sum_i ||y_i - B_i theta_i||^2, sum_half = sum_i floor(n_i/2) (MV: floor(n*P/2)),
 * n_points = sum_i n_i. */
int bfmmm_ssr(bfmmm_engine* e, double* ssr, double* sum_half, double* n_points);
/* sufficient statistics feeding updateNu / updatePhi / updateEta / updateXi (UpdateNu.h:42-63,
 * UpdatePhi.h:44-71, UpdateEta.h:51-81, UpdateXi.h:51-72) with q = K(1+D)(1+M) features ordered
 * f = (k*(1+M) + m')*(1+D) + d'  (m' = 0: mean block, m' = m+1: eigen block m; d' = 0: no
 * covariate, d' = d+1: covariate d):
 *   WtW  q x q   column-major  sum_i w_if w_ig
 *   BtYW P x q   column-major  sum_i w_if B_i' y_i
 * common grid / MV only; H_fg = WtW[f,g] * (B'B).  Ragged grids: see bfmmm_suffstats_ragged. */
int bfmmm_suffstats(bfmmm_engine* e, double* WtW, double* BtYW);
/* ragged grids: additionally Hband[pair][j*P + p] = sum_i w_ia w_ib G_i[p-j][p] for every feature pair
 * a <= b (row-major upper triangle, npairs = q(q+1)/2) and the bw = degree+1 stored diagonals j.
 * BtYW is sum_i w_if B_i'y_i as for the common grid; WtW is returned too. */
int bfmmm_suffstats_ragged(bfmmm_engine* e, double* WtW, double* BtYW, double* Hband);
/* shard geometry: dims = {n, K, P, M, D, model, ragged (0/1), band width} (8 ints) */
int bfmmm_engine_dims(bfmmm_engine* e, int32_t* dims);
/* the data-only counts of updateSigma: sum_i floor(n_i/2) (UpdateSigma.h:49; MV floor(n*P/2), :150)
 * and sum_i n_i, for this shard; no kernel launch */
int bfmmm_counts(bfmmm_engine* e, double* sum_half, double* n_points);
/* Gram matrix B'B (P x P, column-major) of the common basis (identity for MV). */
int bfmmm_get_gram(bfmmm_engine* e, double* G);

/* ---- device RNG ----------------------------------------------------------------------------- */
int bfmmm_seed(bfmmm_engine* e, uint64_t key, uint64_t iteration);

/* ---- multi-GPU: the statistics above are per-shard; these expose the device-side buffers so the
 * caller can all-reduce them in place (NCCL) before reading them back ------------------------- */
/* device pointer + length (doubles) of the engine's statistics buffer, laid out as
 * [sum_log_Z K | n_accept 1 | ssr 1 | ssr_after 1 | WtW q*q | BtYW P*q] */
int bfmmm_stats_buffer_dev(bfmmm_engine* e, double** ptr_dev, int64_t* len);
/* *_async variants: launch only, results stay in the statistics buffer (no host sync). */
int bfmmm_update_z_async(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta);
int bfmmm_update_chi_async(bfmmm_engine* e, double beta);
int bfmmm_ssr_async(bfmmm_engine* e);
int bfmmm_suffstats_async(bfmmm_engine* e);
int bfmmm_read_stats(bfmmm_engine* e, double* out, int64_t len);   /* synchronises */
int bfmmm_clear_ssr_after(bfmmm_engine* e);   /* zero the ssr_after slot once its global sum has been consumed */
int bfmmm_sync(bfmmm_engine* e);
/* CUDA stream the engine launches on (cudaStream_t as void*), for event timing by the caller */
void* bfmmm_stream(bfmmm_engine* e);

#ifdef __cplusplus
}
#endif
#endif
