/*
 * bfmmm_sampler.h -- host loop of the sampler above the engine ABI (include/bfmmm.h).
 *
 * This is the Rcpp-free restatement of the reference's driver loops:
 *   BFMMM_Theta            inst/include/BayesFMMM/BFMMM.h:1253-1298   (Phi,delta,A,gamma,tau,sigma,chi,loglik)
 *   BFMMM_Nu_Z             BFMMM.h:1073-1113                          (Z,pi,alpha3,nu,tau,sigma,loglik)
 *   BFMMM_MTT_warm_start   BFMMM.h:1500-1554                          (Z,pi,alpha3,Phi,delta,A,gamma,nu,tau,sigma,chi)
 * and their MV / covariate-adjusted twins.  Everything that scales with the number of functions
 * runs on the device through the engine; the small prior updates and the P x P Gaussian block
 * draws (updateNu/updatePhi/updateEta/updateXi's pinv/inv + mvnrnd) run here on the host from the
 * device's sufficient statistics, in the reference's sequential block order.
 *
 * In R the patched BFMMM.h drivers would keep their own loop and call the engine directly
 * (INTEGRATION.md); this loop exists so the whole iteration can be tested and benchmarked
 * without R, and it is what bench.py times.
 *
 * Random numbers: global draws come from counter-based Philox streams keyed by
 * (seed, iteration, purpose) so that every rank of a multi-GPU run draws identical globals;
 * with bfmmm_sampler_tape() they are instead popped from a FIFO the caller fills (parity tests).
 */
#ifndef BFMMM_SAMPLER_H
#define BFMMM_SAMPLER_H
#include "bfmmm.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bfmmm_sampler bfmmm_sampler;

/* R defaults: src/UserFunctions.cpp:696-714 */
typedef struct {
  double c[8];             /* Dirichlet prior of pi (length K), default 10 */
  double b;                /* rate of the alpha_3 prior, default 10 */
  double nu_1;             /* gamma shrinkage df, default 3 */
  double alpha1l, alpha2l, beta1l, beta2l;
  double a_Z_PM, a_pi_PM;  /* proposal concentrations, defaults 10000 / 1000 */
  double var_alpha3, var_epsilon1, var_epsilon2;
  double alpha_nu, beta_nu;    /* "alpha", "beta" of updateTau */
  double alpha_eta, beta_eta;
  double alpha_0, beta_0;      /* sigma^2 prior */
} bfmmm_hyper;

void bfmmm_hyper_defaults(bfmmm_hyper* h, int theta_est_defaults);

enum { BFMMM_SWEEP_THETA = 0, BFMMM_SWEEP_NU_Z = 1, BFMMM_SWEEP_FULL = 2 };

/* all-reduce hook for multi-GPU runs: sum `len` doubles at device pointer `buf_dev` over all ranks,
 * ordered on `stream` (cudaStream_t).  NULL => single shard. */
typedef int (*bfmmm_allreduce_fn)(void* ctx, double* buf_dev, int64_t len, void* stream);

/* n_total = number of functions over ALL shards (= engine n on one GPU); Pmat = P x P penalty
 * (column-major) or NULL for the multivariate model's identity prior. */
int bfmmm_sampler_create(bfmmm_engine* e, const bfmmm_hyper* h, int64_t n_total, const double* Pmat,
                         uint64_t seed, bfmmm_sampler** out);
/* sampler without an engine, for exercising the bfmmm_host_update_* functions on the CPU:
 * dims = {n, K, P, M, D, model, ragged (0/1), band width} (8 ints); G = B'B (P x P) or NULL for identity */
int bfmmm_sampler_create_detached(const int32_t* dims, const bfmmm_hyper* h, int64_t n_total, const double* Pmat,
                                  const double* G, double sum_half_total, double n_points_total, uint64_t seed,
                                  bfmmm_sampler** out);
void bfmmm_sampler_destroy(bfmmm_sampler* s);
/* 1 when the sweeps run device-resident: common basis (functional or multivariate) without covariates, no injected
 * draws, and block precisions small enough for one thread block; the global parameters then live in device memory,
 * every update of the sweep is a kernel (csrc/globals_kernels.cu) and the host only reads on demand.  Selected with
 * the environment variable BFMMM_DEVICE_GLOBALS=1 when the sampler is created (default: the host draws the globals,
 * which is faster on one GPU: DESIGN.md section 4). */
int bfmmm_sampler_device_resident(bfmmm_sampler* s);
int bfmmm_sampler_set_allreduce(bfmmm_sampler* s, bfmmm_allreduce_fn fn, void* ctx);
/* Native NCCL hook (csrc/nccl_hook.cu): ncclAllReduce on the engine's stream, no host-language round trip.
 * libnccl is dlopen'ed (libnccl_path, or the library already mapped into the process when NULL).  Rank 0
 * calls bfmmm_nccl_unique_id and distributes the 128 bytes; every rank then calls
 * bfmmm_sampler_enable_nccl (a collective) with its CUDA device current. */
int bfmmm_nccl_unique_id(const char* libnccl_path, char* id_out /* 128 bytes */);
int bfmmm_sampler_enable_nccl(bfmmm_sampler* s, const char* libnccl_path, const char* id /* 128 bytes */, int rank,
                              int world, void** comm_out);
void bfmmm_nccl_destroy(void* comm);
/* One-shot all-reduce over NVLink peer memory (csrc/p2p_hook.cu) for ranks of one node with peer access:
 * every rank calls bfmmm_p2p_create (its mailbox + 64-byte CUDA IPC handle), the handles are all-gathered
 * by the caller, then bfmmm_sampler_enable_p2p maps the peers and installs the hook.  Results are summed in
 * rank order on every rank (bit-identical).  A barrier over all ranks must separate enable from the first
 * sweep and the last sweep from bfmmm_p2p_destroy.  With this hook the engine runs the whole-buffer exchange inside
 * the statistics pass's final reduction and the SSR slot's inside the SSR pass (no kernels of their own), so while it is installed bfmmm_suffstats* on this engine is a
 * collective every rank must enter; that wiring
 * ends with bfmmm_sampler_destroy or a later bfmmm_sampler_set_allreduce: destroy the sampler BEFORE bfmmm_p2p_destroy. */
int bfmmm_p2p_create(int rank, int world, int64_t cap, void** ctx_out, char* handle_out /* 64 bytes */);
int bfmmm_sampler_enable_p2p(bfmmm_sampler* s, void* ctx, const char* handles /* world x 64 bytes */);
void bfmmm_p2p_destroy(void* ctx);
/* ragged grids: the pair cross-Gram band (bfmmm_suffstats_ragged) used by the block draws that follow */
int bfmmm_sampler_set_hband(bfmmm_sampler* s, const double* Hband);
/* totals over all shards of sum_i floor(n_i/2) and sum_i n_i (defaults: this shard's counts scaled by n_total/n) */
int bfmmm_sampler_set_counts(bfmmm_sampler* s, double sum_half_total, double n_points_total);

/* current values, Armadillo layouts; NULL pointers are skipped.
 * nu KxP | Phi KxPxM | pi K | delta KxM | gamma KxPxM | A Kx2 | tau K
 * eta PxDxK | xi K cubes PxDxM | tau_eta KxD | delta_xi KxMxD | gamma_xi K cubes PxDxM | A_xi Kx2xD */
int bfmmm_sampler_set(bfmmm_sampler* s, const double* nu, const double* Phi, const double* sigma_sq,
                      const double* pi, const double* alpha3, const double* delta, const double* gamma,
                      const double* A, const double* tau);
/* NOTE (multi-GPU): with an all-reduce hook installed, asking for `loglik` completes the deferred post-chi SSR
 * exchange, i.e. bfmmm_sampler_get(..., loglik != NULL) is a COLLECTIVE: call it on every rank or on none.
 * With loglik == NULL the call is rank-local. */
int bfmmm_sampler_get(bfmmm_sampler* s, double* nu, double* Phi, double* sigma_sq, double* pi,
                      double* alpha3, double* delta, double* gamma, double* A, double* tau, double* loglik);
int bfmmm_sampler_set_cov(bfmmm_sampler* s, const double* eta, const double* xi, const double* tau_eta,
                          const double* delta_xi, const double* gamma_xi, const double* A_xi);
int bfmmm_sampler_get_cov(bfmmm_sampler* s, double* eta, double* xi, double* tau_eta, double* delta_xi,
                          double* gamma_xi, double* A_xi);

/* one sweep of the chosen driver loop at temperature beta (1 = untempered) */
int bfmmm_sampler_step(bfmmm_sampler* s, int sweep, double beta);
/* n sweeps back to back (the loop bench.py times) */
int bfmmm_sampler_run(bfmmm_sampler* s, int sweep, int n_iter);
/* One tempered transition of BFMMM_MTT_warm_start (BFMMM.h:1556-1651, acceptance
 * CalculateTTAcceptance.h:22-97): 2*N_t tempered sweeps over the geometric ladder
 * (BFMMM.h:1452-1460), accepted with probability min(1, exp(logA)); on rejection every global and
 * the device-resident Z, chi are restored. */
int bfmmm_sampler_tempered_transition(bfmmm_sampler* s, int N_t, double beta_N_t, double* logA, int* accepted);
/* n iterations with the reference's schedule: a tempered transition replaces the sweep whenever
 * the iteration index is a positive multiple of n_temp_trans (0 = never, UserFunctions.cpp:1353-1359) */
int bfmmm_sampler_run_mtt(bfmmm_sampler* s, int n_iter, int n_temp_trans, int N_t, double beta_N_t);
/* write the reference's stored-sample files (include/bfmmm_io.h; BFMMM.h:1680-1746): one batch
 * directory + {Nu,Chi,Pi,alpha_3,A,Delta,Sigma,Tau,Z,Gamma,Phi}{q}.txt every r_stored_iters iterations,
 * keeping slot 0 and every thinning_num-th draw, exactly like BFMMM_MTT_warm_start */
int bfmmm_sampler_record(bfmmm_sampler* s, const char* directory, int r_stored_iters, int thinning_num);
int bfmmm_sampler_batches_written(bfmmm_sampler* s);
/* per-slot (SSR, sigma^2) of the last tempered transition, slots 0..2*N_t */
int bfmmm_sampler_tt_trace(bfmmm_sampler* s, double* ssr, double* sigma, int n);
/* wall-clock split of the sweeps so far, seconds: {host draws, waiting on the device, pushing globals} */
int bfmmm_sampler_profile(bfmmm_sampler* s, double* out3);
int64_t bfmmm_sampler_iteration(bfmmm_sampler* s);
/* acceptance count of the last Z step (summed over shards) */
int64_t bfmmm_sampler_last_accept(bfmmm_sampler* s);

/* positions the global random streams (seed, tick, purpose) used by the bfmmm_host_update_* calls that follow;
 * bfmmm_sampler_step advances the tick itself */
int bfmmm_sampler_set_tick(bfmmm_sampler* s, int64_t tick);
/* injected draws for the parity tests: values are consumed in the reference's call order */
int bfmmm_sampler_tape(bfmmm_sampler* s, const double* values, int64_t n);
int64_t bfmmm_sampler_tape_left(bfmmm_sampler* s);

/* ---- the individual host updates (exposed so tests can check each against the reference) ---- */
int bfmmm_host_update_pi(bfmmm_sampler* s, const double* sum_log_Z);                 /* UpdatePi.h:84-116 */
int bfmmm_host_update_alpha3(bfmmm_sampler* s, const double* sum_log_Z);             /* UpdateAlpha3.h:36-63 */
int bfmmm_host_update_tau(bfmmm_sampler* s);                                         /* UpdateTau.h:18-68 */
int bfmmm_host_update_delta(bfmmm_sampler* s);                                       /* UpdateDelta.h:17-66 */
int bfmmm_host_update_gamma(bfmmm_sampler* s);                                       /* UpdateGamma.h:17-38 */
int bfmmm_host_update_A(bfmmm_sampler* s);                                           /* UpdateA.h:58-135 */
int bfmmm_host_update_tau_eta(bfmmm_sampler* s);                                     /* UpdateTau.h:75-128 */
int bfmmm_host_update_delta_xi(bfmmm_sampler* s);                                    /* UpdateDelta.h:76-125 */
int bfmmm_host_update_gamma_xi(bfmmm_sampler* s);                                    /* UpdateGamma.h:48-72 */
int bfmmm_host_update_A_xi(bfmmm_sampler* s);                                        /* UpdateA.h:137-209 */
/* Gaussian block draws from given sufficient statistics (q x q, P x q, column-major) */
int bfmmm_host_update_phi(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta);
int bfmmm_host_update_nu(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta);
int bfmmm_host_update_eta(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta);
int bfmmm_host_update_xi(bfmmm_sampler* s, const double* WtW, const double* BtYW, double beta);
int bfmmm_host_update_sigma(bfmmm_sampler* s, double ssr, double beta, int tempered);  /* UpdateSigma.h:47-53 */

#ifdef __cplusplus
}
#endif
#endif
