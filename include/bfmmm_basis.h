/*
 * bfmmm_basis.h -- basis and penalty construction of the drivers, host side.
 *
 *   bfmmm_bspline_basis   splines2::BSpline(t, internal_knots, degree, boundary_knots).basis(true)
 *                         as called at BFMMM.h:1188-1196 (clamped, intercept kept, right end closed)
 *   bfmmm_tensor_bspline  BayesFMMM::TensorBSpline, BSplines.h:18-62 (BHDFMMM_* drivers, BFMMM.h:3069)
 *   bfmmm_get_P           BayesFMMM::GetP, BSplines.h:70-120
 *   bfmmm_pmat_rw1        the tridiagonal first-difference penalty built inline at BFMMM.h:1198-1208
 *
 * The high-dimensional functional model (BHDFMMM_Theta_est) is the functional engine with
 * B = bfmmm_tensor_bspline(...) passed as the common-grid basis and P_mat = bfmmm_get_P(...).
 * (For a univariate common grid the engine can also evaluate the basis on the device: cfg.B = NULL.)
 * All matrices are returned in the layout the engine takes: B row-major (point x P), P_mat
 * column-major (symmetric).
 */
#ifndef BFMMM_BASIS_H
#define BFMMM_BASIS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int bfmmm_bspline_basis(const double* t, int64_t n, const double* internal_knots, int n_internal, int degree,
                        double boundary_lo, double boundary_hi, double* B_rowmajor);
/* t: n x dim column-major; boundary: dim x 2 row-major; internal knots of all dimensions back to back */
int bfmmm_tensor_bspline(const double* t, int64_t n, int dim, const int32_t* degree, const double* boundary,
                         const double* internal_knots, const int32_t* n_internal, double* B_rowmajor);
int bfmmm_tensor_P(int dim, const int32_t* degree, const int32_t* n_internal);   /* returns P (product of the per-dimension sizes) */
int bfmmm_get_P(int dim, const int32_t* degree, const int32_t* n_internal, double* P_mat);
int bfmmm_pmat_rw1(int P, double* P_mat);
#ifdef __cplusplus
}
#endif
#endif
