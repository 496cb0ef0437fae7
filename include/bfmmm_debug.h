/* bfmmm_debug.h -- diagnostics entry points of libbfmmm_b200.so used by the parity tests only.
 * They are not part of the drop-in boundary (include/bfmmm.h). */
#ifndef BFMMM_DEBUG_H
#define BFMMM_DEBUG_H
#include "bfmmm.h"
#ifdef __cplusplus
extern "C" {
#endif
/* keep / return the per-function Metropolis log acceptance ratio of bfmmm_update_z */
int bfmmm_debug_enable_acc(bfmmm_engine* e, int on);
int bfmmm_debug_get_acc(bfmmm_engine* e, double* acc /* n */);
/* device-RNG updates that also return the draws they used (to replay them through the oracle) */
int bfmmm_debug_update_z_rng(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM, double beta,
                             double* gam_out /* n x K */, double* u_out /* n */);
int bfmmm_debug_update_chi_rng(bfmmm_engine* e, double beta, double* eps_out /* n x M */);
/* the proposal half of the Z step (z_propose_kernel) on the engine's stream, for kernel timing; fails when the engine
 * runs the Z step as one kernel (ragged grids) */
int bfmmm_debug_z_propose(bfmmm_engine* e, const double* pi, double alpha3, double a_Z_PM);
/* 1 when the next bfmmm_update_chi will draw from the per-function moments the preceding bfmmm_ssr left (common basis,
 * no covariates), 0 when it will make its own pass over the cache */
int bfmmm_debug_moments_valid(bfmmm_engine* e);
/* the projected cache: whitened coefficients (n x P column-major) and orthogonal residual norms */
int bfmmm_debug_get_cache(bfmmm_engine* e, double* Ct, double* rss);
/* the device routines of csrc/fastmath.cuh applied elementwise (host buffers): which = 0 log, 1 reciprocal,
 * 2 sqrt, 3 log-Gamma, 4/5 cosine/sine of the circle point of a 32-bit word, 6 log1p series,
 * 7 uniform from 52 bits, 8/9 the Box-Muller pair of three words derived from x */
int bfmmm_debug_fastmath(int which, const double* x, double* y, int64_t n);
#ifdef __cplusplus
}
#endif
#endif
