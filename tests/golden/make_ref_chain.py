"""Long chains of the REFERENCE's own update functions, as posterior-level golden fixtures.

The reference ships three 150-iteration example chains (inst/test-data/*_trace); they come from the covariate-adjusted
examples of man/BFMMM_warm_start.Rd (X = rnorm(40), not stored) and are far too short to pin a posterior.  This script
instead runs the full warm-start sweep (BFMMM_MTT_warm_start order, BFMMM.h:1500-1554: Z, pi, alpha_3, Phi, delta, A,
gamma, nu, tau, sigma^2, chi) with the reference's OWN functions -- oracle/_ref: Update*.h compiled from
/root/reference over oracle/shim, every random draw injected from NumPy with the law the reference asks R for -- on
the reference's example data (Sim_data.RDS: 40 functions x 100 points; MVSim_data.RDS: 20 x 10), and stores posterior
means, standard deviations, 5 / 95 % quantiles and batch-means Monte-Carlo standard errors of

    sigma^2,  the fitted coefficients  theta_i = Z_i nu + sum_m chi_im Z_i Phi_m  (n x P),  the mean part  Z nu  (n x P),
    the pointwise variances  diag(U_i U_i'), U_i = sum_k Z_ik Phi_k  (n x P),  Z (n x K)  and  nu (K x P).

    python tests/golden/make_ref_chain.py          (build container only: needs /root/reference; ~2 minutes)

tests/test_gpu_sampler.py::test_long_chain_matches_reference_chain compares the engine's chain with these summaries.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402
from tests import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# hyper-parameters: the R defaults of BFMMM_Theta_est (src/UserFunctions.cpp:696-714) with a_Z_PM = a_pi_PM = 1000
HYP = dict(c=10.0, b=10.0, nu_1=3.0, alpha1l=2.0, alpha2l=3.0, beta1l=2.0, beta2l=2.0, a_Z_PM=1000.0, a_pi_PM=1000.0,
           var_alpha3=0.05, var_epsilon1=1.0, var_epsilon2=1.0, alpha_nu=10.0, beta_nu=1.0, alpha_0=1.0, beta_0=1.0)
SWEEPS, BURN, EVERY = 16000, 4000, 8


def batch_mcse(x, nb=20):
    """Monte-Carlo standard error of the mean by batch means (x: draws x ...)."""
    m = (x.shape[0] // nb) * nb
    bm = x[:m].reshape(nb, m // nb, *x.shape[1:]).mean(axis=1)
    return bm.std(axis=0, ddof=1) / np.sqrt(nb)


def initial_state(family, n, K, P, M):
    """The starting point both chains use: the stored example chain's medians for Z, nu, pi (trace_summaries.npz)."""
    S = np.load(os.path.join(HERE, "trace_summaries.npz"))
    rng = np.random.default_rng(4)
    Z = S[f"{family}_Z_med"]; Z = np.asfortranarray(Z / Z.sum(axis=1, keepdims=True))
    return dict(Z=Z, chi=np.asfortranarray(rng.normal(size=(n, M))), nu=np.asfortranarray(S[f"{family}_nu_med"]),
                Phi=np.asfortranarray(0.1 * rng.normal(size=(K, P, M))), pi=np.array(S[f"{family}_pi_med"]), sigma_sq=1.0, alpha3=1.0)


def run_chain(d, init, Pm, mv, seed):
    rng = np.random.default_rng(seed)
    n, K, P, M = d.n, d.K, d.P, d.M
    Z, chi, nu, Phi = init["Z"].copy(), init["chi"].copy(), init["nu"].copy(), init["Phi"].copy()
    pi, sigma_sq, alpha3 = init["pi"].copy(), init["sigma_sq"], init["alpha3"]
    delta = np.ones((K, M), order="F"); gamma = np.ones((K, P, M), order="F"); A = np.ones((K, 2), order="F"); tau = np.ones(K)
    n_half = (n * P) // 2 if mv else int(np.sum((d.off[1:] - d.off[:-1]) // 2))      # UpdateSigma.h:49 / :150
    keep = {k: [] for k in ("sigma", "theta", "fit", "fvar", "Z", "nu")}
    for it in range(SWEEPS):
        st = orc.State(nu=nu, Phi=Phi, Z=Z, chi=chi, sigma_sq=sigma_sq)
        # Z (UpdateMixedMembership.h:131-185): K gammas + 1 uniform per function; non-positive concentrations -> 10
        sh = HYP["a_Z_PM"] * Z
        Z = np.asfortranarray(ref.update_z(d, st, pi, alpha3, HYP["a_Z_PM"], rng.gamma(np.where(sh > 0, sh, 10.0)), rng.uniform(size=n)))
        pi = ref.update_pi(Z, np.full(K, HYP["c"]), alpha3, HYP["a_pi_PM"], pi, rng.gamma(HYP["a_pi_PM"] * pi), rng.uniform())
        alpha3 = ref.update_alpha3(Z, pi, HYP["b"], HYP["var_alpha3"], alpha3, rng.uniform(), rng.uniform())
        st = orc.State(nu=nu, Phi=Phi, Z=Z, chi=chi, sigma_sq=sigma_sq)
        tt = np.asfortranarray(np.cumprod(delta, axis=1))                                   # tilde_tau, BFMMM.h:1254-1259
        Phi = np.asfortranarray(ref.update_phi(d, st, gamma, tt, rng.normal(size=(P, K * M))))
        shp = np.array([[A[k, 0] + P * M / 2.0 if i == 0 else A[k, 1] + P * (M - i) / 2.0 for i in range(M)] for k in range(K)])
        delta = np.asfortranarray(ref.update_delta(Phi, gamma, A, delta, rng.gamma(shp).ravel()))   # order k outer, i inner
        A = np.asfortranarray(ref.update_A(HYP["alpha1l"], HYP["beta1l"], HYP["alpha2l"], HYP["beta2l"], delta, HYP["var_epsilon1"],
                                           HYP["var_epsilon2"], A, rng.uniform(size=K * 2 * 2)))
        gamma = np.asfortranarray(ref.update_gamma(HYP["nu_1"], delta, Phi, rng.gamma((HYP["nu_1"] + 1) / 2, size=K * P * M)))
        st = orc.State(nu=nu, Phi=Phi, Z=Z, chi=chi, sigma_sq=sigma_sq)
        nu = np.asfortranarray(ref.update_nu(d, st, tau, Pm, rng.normal(size=(P, K))))
        tau = ref.update_tau(nu, Pm, HYP["alpha_nu"], HYP["beta_nu"], rng.gamma(HYP["alpha_nu"] + P // 2, size=K), mv=mv)
        st = orc.State(nu=nu, Phi=Phi, Z=Z, chi=chi, sigma_sq=sigma_sq)
        sigma_sq = ref.update_sigma(d, st, HYP["alpha_0"], HYP["beta_0"], rng.gamma(HYP["alpha_0"] + n_half))
        st = orc.State(nu=nu, Phi=Phi, Z=Z, chi=chi, sigma_sq=sigma_sq)
        chi = np.asfortranarray(ref.update_chi(d, st, rng.normal(size=(n, M))))
        if it >= BURN and (it - BURN) % EVERY == 0:
            U = np.einsum("nk,kpm->npm", Z, Phi)
            keep["sigma"].append(sigma_sq); keep["fit"].append(Z @ nu); keep["fvar"].append((U ** 2).sum(axis=2))
            keep["theta"].append(Z @ nu + np.einsum("nm,npm->np", chi, U)); keep["Z"].append(Z.copy()); keep["nu"].append(nu.copy())
    return {k: np.array(v) for k, v in keep.items()}


def summarise(prefix, ch, out):
    for k, x in ch.items():
        out[f"{prefix}_{k}_mean"] = x.mean(axis=0); out[f"{prefix}_{k}_sd"] = x.std(axis=0, ddof=1)
        out[f"{prefix}_{k}_q05"] = np.quantile(x, 0.05, axis=0); out[f"{prefix}_{k}_q95"] = np.quantile(x, 0.95, axis=0)
        out[f"{prefix}_{k}_mcse"] = batch_mcse(x)


def main():
    assert ref.available(), "oracle/_ref could not be built (needs /root/reference)"
    G = np.load(os.path.join(HERE, "sim_inputs.npz"))
    out = {"sweeps": SWEEPS, "burn": BURN, "every": EVERY}
    # functional: Sim_data.RDS, K = 2, cubic basis with internal knots 250 / 500 / 750 (P = 7), M = 3
    y, t = G["sim_y"], G["sim_t"][0]
    n, T, K, P, M = 40, 100, 2, 7, 3
    B = synth.bspline_design(t, [250.0, 500.0, 750.0], 3, (0.0, 1000.0))
    d = orc.Data(n=n, K=K, P=P, M=M, y=y.ravel(), B=np.tile(B, (n, 1)), off=np.arange(n + 1, dtype=np.int64) * T)
    summarise("Functional", run_chain(d, initial_state("Functional", n, K, P, M), orc.pmat_rw1(P), False, 2024), out)
    # multivariate: MVSim_data.RDS 20 x 10, K = 2, M = 2
    y = np.asfortranarray(G["mv_y"])
    n, R, K, M = 20, 10, 2, 2
    d = orc.Data(n=n, K=K, P=R, M=M, y=y, identity_basis=True)
    summarise("Multivariate", run_chain(d, initial_state("Multivariate", n, K, R, M), None, True, 2025), out)
    np.savez_compressed(os.path.join(HERE, "ref_chain_summaries.npz"), **out)
    for fam in ("Functional", "Multivariate"):
        print(fam, "sigma^2 mean", out[fam + "_sigma_mean"], "sd", out[fam + "_sigma_sd"], "mcse", out[fam + "_sigma_mcse"],
              "| theta mcse/sd median", float(np.median(out[fam + "_theta_mcse"] / out[fam + "_theta_sd"])),
              "| fit mcse/sd median", float(np.median(out[fam + "_fit_mcse"] / out[fam + "_fit_sd"])))


if __name__ == "__main__":
    main()
