"""GPU parity: every hot-path entry point of the C ABI against the CPU oracle on the same seeded
inputs and the same injected draws.  FP64 tolerance 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import cases
from tests.gpu_util import engine_for, rel

pytestmark = pytest.mark.gpu
TOL = 1e-10
GPU_CASES = list(cases.CASES)          # common grid, multivariate, covariate-adjusted and ragged grids
COMMON_CASES = [c for c in cases.CASES if cases.CASES[c][0] != "ragged"]
RAGGED_CASES = [c for c in cases.CASES if cases.CASES[c][0] == "ragged"]


def z_ratio_tolerance(Z, gam, acc_o, a_Z_PM):
    """Derived bound on |device - oracle| for the Z step's Metropolis log-ratio.  The reference (and the oracle,
    which follows it term by term) adds 2(K+1) log-Gammas of magnitude lgamma(a z) ~ a log a and 2K products
    (a z - 1) log z (UpdateMixedMembership.h:102-113, Distributions.h:51-61); each summand carries at least half an
    ulp of rounding error, so the reference's own value is only defined to about eps * L, L = the sum of the
    summands' magnitudes (2e5..1e6 at a = 2e4, i.e. ~1e-10 absolute).  The device evaluates the same quantity in a
    cancellation-free closed form (pass_kernels.cuh, z_logratio_closed) good to ~1e-12, so the comparison is held to
    1e-10 relative to the ratio itself plus 4 eps L for the oracle's rounding."""
    from scipy.special import gammaln
    Zs = gam / gam.sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        L = (np.abs(gammaln(a_Z_PM * Z)).sum(axis=1) + np.abs(gammaln(a_Z_PM * Zs)).sum(axis=1)
             + np.abs(gammaln(a_Z_PM * Z.sum(axis=1))) + np.abs(gammaln(a_Z_PM * Zs.sum(axis=1)))
             + (np.abs(a_Z_PM * Z - 1) * np.abs(np.log(Zs))).sum(axis=1)
             + (np.abs(a_Z_PM * Zs - 1) * np.abs(np.log(Z))).sum(axis=1))
    return TOL * (1.0 + np.abs(acc_o)) + 4 * np.finfo(float).eps * L


@pytest.mark.parametrize("name", GPU_CASES)
@pytest.mark.parametrize("beta", [1.0, 0.6])
def test_update_z(name, beta):
    s, d, st, eng = engine_for(name)
    dr = cases.draws(name, s)
    eng.debug_enable_acc(True)
    Zo, acc_o, took = orc.update_z(d, st, s["pi"], 1.3, cases.A_Z_PM, dr["gam"], dr["u"], beta)
    slz, nacc = eng.update_z(s["pi"], 1.3, cases.A_Z_PM, beta, gam=dr["gam"], u=dr["u"])
    Zg, _ = eng.get_state(chi=False)
    acc_g = eng.debug_get_acc()
    # a proposal coordinate that underflows to 0 makes the reference's ratio NaN (=> reject): same here
    assert np.array_equal(np.isnan(acc_g), np.isnan(acc_o))
    fin = ~np.isnan(acc_o)
    tol = z_ratio_tolerance(s["Z"], dr["gam"], acc_o, cases.A_Z_PM)
    assert np.all(np.abs(acc_g[fin] - acc_o[fin]) <= tol[fin]), np.max(np.abs(acc_g[fin] - acc_o[fin]) / tol[fin])
    margin = np.abs(np.log(dr["u"]) - acc_o)
    decided = ~(margin <= 1e-8)      # knife-edge decisions may legitimately differ
    assert np.array_equal(Zg[decided], Zo[decided])
    assert nacc == int(took.sum()) or not decided.all()
    assert rel(slz, np.log(Zo).sum(axis=0)) < TOL
    eng.close()


@pytest.mark.parametrize("name", GPU_CASES)
@pytest.mark.parametrize("beta", [1.0, 0.6])
@pytest.mark.parametrize("after_ssr", [False, True])
def test_update_chi_and_ssr_after(name, beta, after_ssr):
    """updateChi on its own (chi_kernel's pass over the cache) and right after the SSR pass of updateSigma, the order of
    the sweep: on a common basis without covariates the SSR pass leaves the per-function moments and the chi step draws
    from them without a second pass (moments_kernels.cu).  Both must give the oracle's chi."""
    s, d, st, eng = engine_for(name)
    dr = cases.draws(name, s)
    chi_o = orc.update_chi(d, st, dr["eps"], beta)
    if after_ssr:
        assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
        assert eng.debug_moments_valid() == (cases.CASES[name][0] in ("common", "mv", "hd") and d.D == 0 and d.M >= 1)
    ssr_after = eng.update_chi(beta, eps=dr["eps"])
    _, chi_g = eng.get_state(Z=False)
    assert rel(chi_g, chi_o) < TOL
    st2 = orc.State(nu=st.nu, Phi=st.Phi, Z=st.Z, chi=chi_o, sigma_sq=st.sigma_sq, eta=st.eta, xi=st.xi)
    assert rel(ssr_after, orc.ssr(d, st2)[0]) < TOL
    eng.close()


@pytest.mark.parametrize("name", GPU_CASES)
def test_ssr_sigma_loglik(name):
    s, d, st, eng = engine_for(name)
    ssr_o, half_o, npts_o = orc.ssr(d, st)
    ssr_g, half_g, npts_g = eng.ssr()
    assert rel(ssr_g, ssr_o) < TOL
    assert half_g == half_o and npts_g == npts_o
    # sigma^2 draw and log-likelihood assembled on the host from the device statistic
    g = 37.5
    sig_o, a_o, b_o = orc.update_sigma(d, st, 1.0, 1.0, g)
    sig_g = 1.0 / ((1.0 / (0.5 * ssr_g + 1.0)) * g)
    assert rel(sig_g, sig_o) < TOL
    ll_o = orc.loglik(d, st)
    if d.identity_basis:
        ll_g = -(d.n * (d.P // 2)) * np.log(2 * np.pi * st.sigma_sq) - ssr_g / (2 * st.sigma_sq)
    else:
        ll_g = -npts_g * (0.5 * np.log(2 * np.pi) + 0.5 * np.log(st.sigma_sq)) - ssr_g / (2 * st.sigma_sq)
    assert rel(ll_g, ll_o) < TOL
    eng.close()


def _features(s, d):
    """W (n x q) in the engine's feature order f = ((k*(1+M) + m')*(1+D) + d')."""
    n, K, M, D = d.n, d.K, d.M, d.D
    cols = []
    for k in range(K):
        for mm in range(M + 1):
            for dd in range(D + 1):
                w = s["Z"][:, k].copy()
                if mm:
                    w = w * s["chi"][:, mm - 1]
                if dd:
                    w = w * s["X"][:, dd - 1]
                cols.append(w)
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("name", COMMON_CASES)
def test_suffstats(name):
    s, d, st, eng = engine_for(name)
    W = _features(s, d)
    WtW, BtYW = eng.suffstats()
    if d.identity_basis:
        BtY = np.asarray(s["y"])                    # n x P
    else:
        BtY = s["y"] @ s["B"]                       # (n x T)(T x P)
    assert rel(WtW, W.T @ W) < TOL
    assert rel(BtYW, BtY.T @ W) < TOL
    G = eng.gram()
    Gref = np.eye(d.P) if d.identity_basis else s["B"].T @ s["B"]
    assert rel(G, Gref) < 1e-13
    eng.close()


def test_device_bspline_basis_matches_oracle():
    s, d, st, eng = engine_for("F_common", device_basis=True)
    B = eng.basis(s["T"])
    Bo = orc.bspline_basis(s["t"], s["internal_knots"], s["degree"], s["boundary"])
    assert np.array_equal(B, Bo)
    assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
    eng.close()


def test_projection_cache_identity():
    """||y_i - B theta||^2 == rss_i + ||c~_i - L' theta||^2 for arbitrary theta (the identity every
    per-iteration kernel relies on)."""
    s, d, st, eng = engine_for("F_common")
    Ct, rss = eng.debug_get_cache()
    B = s["B"]
    L = np.linalg.cholesky(B.T @ B)
    rng = np.random.default_rng(5)
    th = rng.normal(size=(d.n, d.P))
    direct = ((s["y"] - th @ B.T) ** 2).sum(axis=1)
    proj = rss + ((Ct - th @ L) ** 2).sum(axis=1)
    assert rel(proj, direct) < 1e-12
    eng.close()


@pytest.mark.parametrize("name", ["F_common", "MV", "F_cov", "F_ragged"] + cases.BASELINE_CASES)
def test_device_rng_replay(name):
    """Device-RNG mode: replaying the draws the kernels generated through the oracle reproduces the
    device result, and the draws have the right law."""
    s, d, st, eng = engine_for(name)
    eng.seed(1234, 7)
    eng.debug_enable_acc(True)
    gam, u = eng.debug_update_z_rng(s["pi"], 1.3, cases.A_Z_PM)
    Zg, _ = eng.get_state(chi=False)
    Zo, acc_o, _ = orc.update_z(d, st, s["pi"], 1.3, cases.A_Z_PM, gam, u)
    # device-RNG mode takes log z* - log z from the proposal's own pieces (no logarithm of a membership): the
    # ratio it accepted with equals the oracle's on the replayed draws
    acc_g = eng.debug_get_acc()
    fin = ~np.isnan(acc_o)
    assert np.all(np.abs(acc_g[fin] - acc_o[fin]) <= z_ratio_tolerance(s["Z"], gam, acc_o, cases.A_Z_PM)[fin])
    decided = np.abs(np.log(u) - acc_o) > 1e-8
    # same accept/reject decision everywhere; the device normalises with a reciprocal (last-bit differences)
    took_g, took_o = np.any(Zg != s["Z"], axis=1), np.any(Zo != s["Z"], axis=1)
    assert np.array_equal(took_g[decided], took_o[decided])
    assert np.allclose(Zg[decided], Zo[decided], rtol=1e-15, atol=0)
    assert np.all((u > 0) & (u < 1)) and np.all(gam >= 0)     # tiny shapes may underflow to 0, as in R
    big = cases.A_Z_PM * s["Z"] > 5
    zscore = ((gam - cases.A_Z_PM * s["Z"]) / np.sqrt(cases.A_Z_PM * s["Z"]))[big]
    assert abs(zscore.mean()) < 5 / np.sqrt(zscore.size) and 0.7 < zscore.std() < 1.3
    # chi with device normals
    eng.set_state(s["Z"], s["chi"])
    eps = eng.debug_update_chi_rng()
    _, chi_g = eng.get_state(Z=False)
    assert rel(chi_g, orc.update_chi(d, st, eps)) < TOL
    assert abs(eps.mean()) < 5 / np.sqrt(eps.size)
    # the same step drawn from the moments the SSR pass leaves (the sweep's order): same normals, same chi
    eng.set_state(s["Z"], s["chi"])
    eng.ssr()
    eps_m = eng.debug_update_chi_rng()
    _, chi_m = eng.get_state(Z=False)
    assert np.array_equal(eps_m, eps) and rel(chi_m, chi_g) < TOL
    # determinism and shard-independence: same seed -> same draws; a shard starting at global
    # offset 8 reproduces rows 8.. of the full run
    eng.set_state(s["Z"], s["chi"])
    gam2, u2 = eng.debug_update_z_rng(s["pi"], 1.3, cases.A_Z_PM)
    assert np.array_equal(gam, gam2) and np.array_equal(u, u2)
    eng.close()


@pytest.mark.parametrize("name", RAGGED_CASES)
@pytest.mark.parametrize("device_basis", [False, True])
def test_suffstats_ragged(name, device_basis):
    """Ragged grids: pair cross-Gram band sum_i w_ia w_ib G_i and sum_i w_if B_i'y_i (DMMA kernels) against
    a per-function NumPy evaluation; both the user-supplied-rows and the device-spline create paths."""
    from tests.test_host_updates import _ragged_stats
    s, d, st, eng = engine_for(name, device_basis=device_basis)
    dims = eng.dims()
    assert dims[6] == 1 and dims[7] == 4
    WtW, BtYW, Hb = eng.suffstats_ragged(dims[7])
    WtW_o, BtYW_o, Hb_o = _ragged_stats(s, d)
    assert rel(WtW, WtW_o) < TOL and rel(BtYW, BtYW_o) < TOL and rel(Hb, Hb_o) < TOL
    assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
    eng.close()


@pytest.mark.parametrize("name", RAGGED_CASES)
def test_ragged_sweep_matches_oracle_updates(name):
    """One host block draw through the engine's ragged statistics equals the oracle's per-point loops."""
    import bayesfmmm_b200 as bf
    s, d, st, eng = engine_for(name)
    dr = cases.draws(name, s)
    smp = bf.Sampler(eng, hyper=bf.default_hyper(True), n_total=d.n, Pmat=orc.pmat_rw1(d.P), seed=3)
    par = s["par"]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=par["sigma_sq"], tau=dr["tau"])
    if d.D:
        smp.set_cov(eta=par["eta"], xi=par["xi"], tau_eta=dr["tau_eta"])
    WtW, BtYW, Hb = eng.suffstats_ragged(4)
    smp.set_hband(Hb)
    smp.tape(dr["z_nu"].ravel(order="F")); smp.host_update("nu", WtW, BtYW, 1.0)
    # measured 3e-15 .. 4e-14 (B200; the C4 shape's 666 pair bands are the largest)
    assert rel(smp.get()["nu"], orc.update_nu(d, st, dr["tau"], orc.pmat_rw1(d.P), dr["z_nu"])) < TOL
    smp.close(); eng.close()


def test_device_rng_distributions():
    """The device generators (Philox4x32-10 + ziggurat normal + Marsaglia-Tsang gamma) have the right
    laws: Kolmogorov-Smirnov against SciPy on 3e5 normals and on gammas of small, moderate and large
    shape, plus the log of the gamma variate that the Z step reuses."""
    import scipy.stats as sst
    import bayesfmmm_b200 as bf
    from bayesfmmm_b200.engine import MULTIVARIATE
    n, K, M, P = 100_000, 3, 3, 4
    rng = np.random.default_rng(0)
    a_Z = 100.0
    Z = np.empty((n, K)); Z[:, 0] = 0.004; Z[:, 1] = 0.03; Z[:, 2] = 1 - Z[:, 0] - Z[:, 1]   # shapes 0.4, 3, 96.6
    eng = bf.Engine(model=MULTIVARIATE, n=n, K=K, P=P, M=M, y=rng.normal(size=(n, P)))
    eng.set_state(Z, rng.normal(size=(n, M)))
    eng.set_globals(rng.normal(size=(K, P)), 0.1 * rng.normal(size=(K, P, M)), 1.0)
    eng.seed(99, 3)
    eps = eng.debug_update_chi_rng()
    x = eps.ravel()
    assert sst.kstest(x, "norm").pvalue > 1e-3
    assert abs(x.mean()) < 4 / np.sqrt(x.size) and abs(x.var() - 1) < 0.01
    assert abs(sst.kurtosis(x)) < 0.05 and abs(sst.skew(x)) < 0.02
    assert abs(np.mean(np.abs(x) > 3.5) / (2 * sst.norm.sf(3.5)) - 1) < 0.25       # the ziggurat tail
    # independence across functions / coordinates
    assert abs(np.corrcoef(eps[:, 0], eps[:, 1])[0, 1]) < 0.01 and abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 0.01
    eng.set_state(Z, None)
    gam, u = eng.debug_update_z_rng(np.ones(K) / K, 1.0, a_Z)
    assert sst.kstest(u, "uniform").pvalue > 1e-3
    for k in range(K):
        shape = a_Z * Z[0, k]
        assert sst.kstest(gam[:, k], "gamma", args=(shape,)).pvalue > 1e-3, (k, shape)
    eng.close()


@pytest.mark.parametrize("name", GPU_CASES)
def test_marginal_loglik_and_cpo(name):
    """Post-processing on the device: the per-function marginal log-likelihood (chi integrated out, through the
    M x M Woodbury form) and the CPO accumulated over a short stored chain equal the oracle's dense
    n_i x n_i evaluation of calcLikelihoodCPO (CalculateLikelihood.h:344-385)."""
    s, d, st, eng = engine_for(name)
    states = cases.stored_iterations(name, st)
    eng.cpo_reset()
    L = []
    for x in states:
        eng.set_state(x.Z, x.chi)
        eng.set_globals(x.nu, x.Phi, x.sigma_sq, eta=x.eta, xi=x.xi)
        lg = eng.marginal_loglik()
        if not d.identity_basis:
            lo = orc.marginal_loglik(d, x)
            assert np.max(np.abs(lg - lo) / np.abs(lo)) < TOL
        L.append(lg)
        eng.cpo_accumulate()
    cpo = eng.cpo_get()
    assert np.max(np.abs(cpo - orc.cpo(np.stack(L))) / np.abs(cpo)) < 1e-12
    assert np.allclose(eng.cpo_get(log_scale=False), np.exp(cpo), rtol=1e-12)
    eng.close()
