// globals_dev.h -- arguments of the device-side global updates (globals_kernels.cu), shared with host_sampler.cu
#pragma once
#include <cuda_runtime.h>

#include "globals_core.cuh"

namespace bf {

struct DrawArgs {
  GlobalsView g;                 // device pointers
  int Pc, P4, QS, q;             // rows of the projected cache, padded rows / feature stride of glob, features K (M + 1)
  int hbG, hbL, hbmax;           // half bandwidths: basis Gram, whitening factor, max over the blocks' precisions
  const double* G;               // P x P column-major (device) or nullptr (identity model)
  const double* L;               // P x Pc column-major whitening factor (device) or nullptr
  const double* stats;           // engine statistics buffer: [sum log Z (K) | accepts | ssr | ssr_after | W'W | C~'W]
  double* glob;                  // P4 x QS whitened coefficients staged by the pass kernels
  const double* sigma_dev;
  double beta;
  int do_phi, do_nu;
  StreamRng rng;
  int* err;                      // set to 1 when a precision is not positive definite
  long long* clk;                // optional: clock64() at the phase boundaries of draw_blocks_kernel (tuning), or nullptr
};
struct SigmaPiArgs {
  GlobalsView g;
  const double* stats;
  double* sigma_dev;
  double shape, scale_ssr;       // sigma^2 = 1 / (Gamma(shape) / (scale_ssr * SSR + beta_0))
  int do_pi;
  StreamRng rng;
};
struct PriorsArgs {
  GlobalsView g;
  int do_phi;
  StreamRng rng;
};
size_t draw_blocks_smem(const DrawArgs& a);
int launch_draw_blocks(const DrawArgs& a, cudaStream_t s);
int launch_sigma(const SigmaPiArgs& a, cudaStream_t s);
int launch_pi_alpha(const SigmaPiArgs& a, cudaStream_t s);
int launch_priors(const PriorsArgs& a, cudaStream_t s);

}  // namespace bf
