"""Host side of the product (bayesfmmm_b200/csrc/host_sampler.cu): every host update of the driver
loop against the REFERENCE's own function (oracle/_ref) on the same injected draws, and the Gaussian
block draws computed from sufficient statistics against the oracle's direct per-point evaluation.
Runs on the CPU: a detached Sampler needs no GPU (it only exercises bfmmm_host_update_*)."""
import numpy as np
import pytest

import bayesfmmm_b200 as bf
from oracle import oracle as orc
from oracle import ref
from tests import cases

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def detached(K, P, M, D=0, n=50, mv=False, G=None, Pmat=None, **hy):
    h = bf.default_hyper(True, **hy)
    return bf.Sampler(hyper=h, n_total=n, Pmat=Pmat, dims=(n, K, P, M, D, 1 if mv else 0), G=G,
                      sum_half_total=n * 7.0, n_points_total=n * 15.0)


def _state(rng, K, P, M):
    return dict(nu=np.asfortranarray(rng.normal(size=(K, P))), Phi=np.asfortranarray(rng.normal(size=(K, P, M))),
                delta=np.asfortranarray(rng.gamma(2, 1, (K, M)) + 0.2), gamma=np.asfortranarray(rng.gamma(2, 1, (K, P, M)) + 0.2),
                A=np.asfortranarray(rng.gamma(2, 1, (K, 2)) + 0.2), tau=rng.gamma(2, 1, K) + 0.2,
                pi=rng.dirichlet(np.ones(K) * 3))


@needs_ref
@pytest.mark.parametrize("K,P,M", [(3, 8, 3), (2, 7, 2), (4, 9, 5)])
def test_prior_updates_match_reference(K, P, M):
    rng = np.random.default_rng(K * 100 + P * 10 + M)
    n = 60
    st = _state(rng, K, P, M)
    Z = rng.dirichlet(np.ones(K) * 2, size=n)
    slz = np.log(Z).sum(axis=0)
    Pm = orc.pmat_rw1(P)
    s = detached(K, P, M, n=n, Pmat=Pm)
    s.set(alpha3=1.7, sigma_sq=0.3, **st)
    h = s.hyper
    # pi
    gam = rng.gamma(h.a_pi_PM * st["pi"]); u = rng.uniform()
    s.tape(np.concatenate([gam, [u]])); s.host_update("pi", slz)
    pi_ref = ref.update_pi(Z, np.array(h.c[:K]), 1.7, h.a_pi_PM, st["pi"], gam, u)
    assert rel(s.get()["pi"], pi_ref) < RTOL and s.tape_left() == 0
    # alpha3 (state now has the possibly-updated pi)
    pi_now = s.get()["pi"]
    for up, ua in [(0.3, 0.9), (0.8, 1e-9), (0.55, 0.5)]:
        a3_in = s.get()["alpha3"]
        s.tape([up, ua]); s.host_update("alpha3", slz)
        a3_ref = ref.update_alpha3(Z, pi_now, h.b, h.var_alpha3, a3_in, up, ua)
        assert abs(s.get()["alpha3"] - a3_ref) <= 1e-12 * max(1.0, abs(a3_ref))
    # tau
    g = rng.gamma(h.alpha_nu + P // 2, size=K)
    s.tape(g); s.host_update("tau")
    assert rel(s.get()["tau"], ref.update_tau(st["nu"], Pm, h.alpha_nu, h.beta_nu, g)) < RTOL
    # delta
    g = rng.gamma(5.0, size=K * M)
    s.tape(g); s.host_update("delta")
    d_ref = ref.update_delta(st["Phi"], st["gamma"], st["A"], st["delta"], g)
    assert rel(s.get()["delta"], d_ref) < RTOL
    # A (uses the new delta)
    us = rng.uniform(size=K * 2 * 2)
    s.tape(us); s.host_update("A")
    A_ref = ref.update_A(h.alpha1l, h.beta1l, h.alpha2l, h.beta2l, d_ref, h.var_epsilon1, h.var_epsilon2, st["A"], us)
    assert rel(s.get()["A"], A_ref) < RTOL
    # gamma
    g = rng.gamma((h.nu_1 + 1) / 2, size=K * P * M)
    s.tape(g); s.host_update("gamma")
    assert rel(s.get()["gamma"], ref.update_gamma(h.nu_1, d_ref, st["Phi"], g)) < RTOL
    s.close()


@needs_ref
def test_mv_tau_and_cov_priors_match_reference():
    rng = np.random.default_rng(77)
    K, P, M, D, n = 3, 6, 2, 2, 40
    st = _state(rng, K, P, M)
    s = detached(K, P, M, D=D, n=n, mv=True)
    s.set(alpha3=1.2, sigma_sq=0.5, **st)
    h = s.hyper
    g = rng.gamma(h.alpha_nu + P // 2, size=K)
    s.tape(g); s.host_update("tau")
    assert rel(s.get()["tau"], ref.update_tau(st["nu"], None, h.alpha_nu, h.beta_nu, g, mv=True)) < RTOL
    eta = np.asfortranarray(rng.normal(size=(P, D, K))); xi = rng.normal(size=(K, P, D, M))
    dxi = np.asfortranarray(rng.gamma(2, 1, (K, M, D)) + 0.2); gxi = rng.gamma(2, 1, (K, P, D, M)) + 0.2
    Axi = np.asfortranarray(rng.gamma(2, 1, (K, 2, D)) + 0.2)
    s.set_cov(eta=eta, xi=xi, tau_eta=np.ones((K, D)), delta_xi=dxi, gamma_xi=gxi, A_xi=Axi)
    g = rng.gamma(h.alpha_eta + P // 2, size=K * D)
    s.tape(g); s.host_update("tau_eta")
    assert rel(s.get_cov()["tau_eta"], ref.update_tau_eta(eta, None, h.alpha_eta, h.beta_eta, g, mv=True)) < RTOL
    g = rng.gamma(4.0, size=K * M * D)
    s.tape(g); s.host_update("delta_xi")
    dref = ref.update_delta_xi(xi, gxi, Axi, dxi, g)
    assert rel(s.get_cov()["delta_xi"], dref) < RTOL
    us = rng.uniform(size=K * 2 * D * 2)
    s.tape(us); s.host_update("A_xi")
    assert rel(s.get_cov()["A_xi"], ref.update_A_xi(h.alpha1l, h.beta1l, h.alpha2l, h.beta2l, dref, h.var_epsilon1,
                                                   h.var_epsilon2, Axi, us)) < RTOL
    g = rng.gamma((h.nu_1 + 1) / 2, size=K * P * D * M)
    s.tape(g); s.host_update("gamma_xi")
    assert rel(s.get_cov()["gamma_xi"], ref.update_gamma_xi(h.nu_1, dref, xi, g)) < RTOL
    # functional tau_eta with the penalty matrix
    Pm = orc.pmat_rw1(P)
    s2 = detached(K, P, M, D=D, n=n, Pmat=Pm)
    s2.set(alpha3=1.2, sigma_sq=0.5, **st)
    s2.set_cov(eta=eta, xi=xi)
    g = rng.gamma(h.alpha_eta + P // 2, size=K * D)
    s2.tape(g); s2.host_update("tau_eta")
    assert rel(s2.get_cov()["tau_eta"], ref.update_tau_eta(eta, Pm, h.alpha_eta, h.beta_eta, g)) < RTOL
    s.close(); s2.close()


def _features(s, d):
    cols = []
    for k in range(d.K):
        for mm in range(d.M + 1):
            for dd in range(d.D + 1):
                w = s["Z"][:, k].copy()
                if mm:
                    w = w * s["chi"][:, mm - 1]
                if dd:
                    w = w * s["X"][:, dd - 1]
                cols.append(w)
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("name", ["F_common", "F_common_K2M2", "MV", "F_cov", "MV_cov", "F_common_P100", "MV_R100"])
@pytest.mark.parametrize("beta", [1.0, 0.6])
def test_block_draws_from_sufficient_statistics(name, beta):
    """nu / Phi / eta / xi drawn on the host from (W'W, B'Y'W) equal the oracle's per-point loops
    (which are pinned to the reference) for the same normals: the one-pass statistics carry the
    reference's sequential block semantics exactly."""
    s, d, st = cases.build(name)
    dr = cases.draws(name, s)
    W = _features(s, d)
    mv = d.identity_basis
    BtY = np.asarray(s["y"]) if mv else s["y"] @ s["B"]
    WtW, BtYW = W.T @ W, BtY.T @ W
    G = None if mv else s["B"].T @ s["B"]
    Pm = None if mv else orc.pmat_rw1(d.P)
    K, P, M, D = d.K, d.P, d.M, d.D
    par = s["par"]
    smp = detached(K, P, M, D=D, n=d.n, mv=mv, G=G, Pmat=Pm)
    delta = np.ones((K, M)); delta[:, 0] = dr["tilde_tau"][:, 0]
    for m in range(1, M):
        delta[:, m] = dr["tilde_tau"][:, m] / dr["tilde_tau"][:, m - 1]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=par["sigma_sq"], tau=dr["tau"], gamma=dr["gamma"], delta=delta)
    if D:
        dxi = np.ones((K, M, D)); dxi[:, 0, :] = dr["tilde_tau_xi"][:, 0, :]
        for m in range(1, M):
            dxi[:, m, :] = dr["tilde_tau_xi"][:, m, :] / dr["tilde_tau_xi"][:, m - 1, :]
        smp.set_cov(eta=par["eta"], xi=par["xi"], tau_eta=dr["tau_eta"], gamma_xi=dr["gamma_xi"], delta_xi=dxi)
    tol = 2e-9
    smp.tape(dr["z_phi"].ravel(order="F")); smp.host_update("phi", WtW, BtYW, beta)
    assert rel(smp.get()["Phi"], orc.update_phi(d, st, dr["gamma"], dr["tilde_tau"], dr["z_phi"], beta)) < tol
    smp.set(Phi=par["Phi"])
    smp.tape(dr["z_nu"].ravel(order="F")); smp.host_update("nu", WtW, BtYW, beta)
    assert rel(smp.get()["nu"], orc.update_nu(d, st, dr["tau"], Pm, dr["z_nu"], beta)) < tol
    smp.set(nu=par["nu"])
    if D:
        smp.tape(dr["z_eta"].ravel(order="F")); smp.host_update("eta", WtW, BtYW, beta)
        assert rel(smp.get_cov()["eta"], orc.update_eta(d, st, dr["tau_eta"], Pm, dr["z_eta"], beta)) < tol
        smp.set_cov(eta=par["eta"])
        smp.tape(dr["z_xi"].ravel(order="F")); smp.host_update("xi", WtW, BtYW, beta)
        assert rel(smp.get_cov()["xi"], orc.update_xi(d, st, dr["gamma_xi"], dr["tilde_tau_xi"], dr["z_xi"], beta)) < tol
    assert smp.tape_left() == 0
    # sigma^2 draw from a given SSR
    smp.tape([41.5]); smp.host_update("sigma", 3.25, beta, beta != 1.0)
    if beta == 1.0:
        expect = 1.0 / ((1.0 / (0.5 * 3.25 + 1.0)) * 41.5)
    else:
        expect = 1.0 / ((1.0 / ((beta / 2) * 3.25 + 1.0)) * 41.5)
    assert abs(smp.get()["sigma_sq"] - expect) < 1e-14
    smp.close()


def _ragged_stats(s, d, bw=4):
    """Pair cross-Gram band and B'Y'W of a ragged data set, computed per function with NumPy."""
    W = _features(s, d)
    q, P, n = W.shape[1], d.P, d.n
    off = s["off"]
    npairs = q * (q + 1) // 2
    Hb = np.zeros((npairs, bw * P)); BtYW = np.zeros((P, q))
    for i in range(n):
        Bi = s["B"][off[i]:off[i + 1]]; yi = s["y"][off[i]:off[i + 1]]
        Gi = Bi.T @ Bi
        band = np.zeros(bw * P)
        for j in range(bw):
            for p in range(j, P):
                band[j * P + p] = Gi[p - j, p]
        assert np.allclose(np.triu(Gi, bw), 0)             # cubic B-splines: bandwidth 4
        pr = 0
        for a in range(q):
            for b in range(a, q):
                Hb[pr] += W[i, a] * W[i, b] * band
                pr += 1
        BtYW += np.outer(Bi.T @ yi, W[i])
    return W.T @ W, BtYW, Hb


@pytest.mark.parametrize("name", ["F_ragged", "F_cov_ragged"])
@pytest.mark.parametrize("beta", [1.0, 0.6])
def test_block_draws_ragged_grids(name, beta):
    """Ragged grids: the per-pair banded cross-Gram carries the reference's block semantics."""
    s, d, st = cases.build(name)
    dr = cases.draws(name, s)
    WtW, BtYW, Hb = _ragged_stats(s, d)
    K, P, M, D = d.K, d.P, d.M, d.D
    Pm = orc.pmat_rw1(P)
    par = s["par"]
    smp = bf.Sampler(hyper=bf.default_hyper(True), n_total=d.n, Pmat=Pm, dims=(d.n, K, P, M, D, 0, 1, 4),
                     sum_half_total=1.0, n_points_total=1.0)
    delta = np.ones((K, M)); delta[:, 0] = dr["tilde_tau"][:, 0]
    for m in range(1, M):
        delta[:, m] = dr["tilde_tau"][:, m] / dr["tilde_tau"][:, m - 1]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=par["sigma_sq"], tau=dr["tau"], gamma=dr["gamma"], delta=delta)
    if D:
        dxi = np.ones((K, M, D)); dxi[:, 0, :] = dr["tilde_tau_xi"][:, 0, :]
        for m in range(1, M):
            dxi[:, m, :] = dr["tilde_tau_xi"][:, m, :] / dr["tilde_tau_xi"][:, m - 1, :]
        smp.set_cov(eta=par["eta"], xi=par["xi"], tau_eta=dr["tau_eta"], gamma_xi=dr["gamma_xi"], delta_xi=dxi)
    smp.set_hband(Hb)
    tol = 2e-9
    smp.tape(dr["z_phi"].ravel(order="F")); smp.host_update("phi", WtW, BtYW, beta)
    assert rel(smp.get()["Phi"], orc.update_phi(d, st, dr["gamma"], dr["tilde_tau"], dr["z_phi"], beta)) < tol
    smp.set(Phi=par["Phi"])
    smp.tape(dr["z_nu"].ravel(order="F")); smp.host_update("nu", WtW, BtYW, beta)
    assert rel(smp.get()["nu"], orc.update_nu(d, st, dr["tau"], Pm, dr["z_nu"], beta)) < tol
    smp.set(nu=par["nu"])
    if D:
        smp.tape(dr["z_eta"].ravel(order="F")); smp.host_update("eta", WtW, BtYW, beta)
        assert rel(smp.get_cov()["eta"], orc.update_eta(d, st, dr["tau_eta"], Pm, dr["z_eta"], beta)) < tol
        smp.set_cov(eta=par["eta"])
        smp.tape(dr["z_xi"].ravel(order="F")); smp.host_update("xi", WtW, BtYW, beta)
        assert rel(smp.get_cov()["xi"], orc.update_xi(d, st, dr["gamma_xi"], dr["tilde_tau_xi"], dr["z_xi"], beta)) < tol
    smp.close()
