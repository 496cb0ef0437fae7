/* bfmmm_post.h -- post-processing reductions over stored draws.
 *
 * The reference's credible-interval functions read the stored batches back and end in arma::quantile per element:
 *   ZCI      src/PostProcessing.cpp:3505-3592   (n x K elements)
 *   SigmaCI  src/PostProcessing.cpp:3435-3480   (1 element)
 *   FMeanCI  src/PostProcessing.cpp:99-480      (T elements; pointwise or simultaneous band)
 *   FCovCI   src/PostProcessing.cpp:1781-2300   (T1 x T2 elements)
 * bfmmm_quantiles is the element-parallel part on the device; bayesfmmm_b200/post.py assembles the four functions
 * on top of it with the reference's argument lists.  (The conditional predictive ordinates are in bfmmm.h:
 * bfmmm_cpo_reset / _accumulate / _get.)
 */
#ifndef BFMMM_POST_H
#define BFMMM_POST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* draws: S x R doubles, draw-major (element r of draw s at draws[s * R + r]); probs: np probabilities;
 * out: np x R (quantile p of element r at out[p * R + r]).  Quantile definition: Armadillo's (Hyndman-Fan type 5).
 * Host pointers; the sort runs on CUDA device `device`.  Returns 0, or 1 with bfmmm_last_error() set. */
int bfmmm_quantiles(const double* draws, int64_t S, int64_t R, const double* probs, int np, double* out, int device);
#ifdef __cplusplus
}
#endif
#endif
