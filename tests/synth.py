"""Seeded synthetic inputs shaped like the reference's unit tests (src/test-Sigma.cpp:8-82,
src/test-Chi.cpp:8-89) and SURVEY.md section 8(d).  Shared by CPU and GPU tests and bench.py."""
import numpy as np
from scipy.interpolate import BSpline


def clamped_knots(internal, degree, boundary):
    return np.concatenate([[boundary[0]] * (degree + 1), np.asarray(internal, float), [boundary[1]] * (degree + 1)])


def bspline_design(t, internal, degree=3, boundary=(0.0, 1000.0)):
    """Dense clamped B-spline design matrix (SciPy; reproduces the reference goldens)."""
    kn = clamped_knots(internal, degree, boundary)
    return BSpline.design_matrix(np.asarray(t, float), kn, degree, extrapolate=False).toarray()


def equispaced_internal(P, degree=3, boundary=(0.0, 1000.0)):
    n_ik = P - degree - 1
    return np.linspace(boundary[0], boundary[1], n_ik + 2)[1:-1]


def make_params(rng, K, P, M, D=0, sigma_sq=0.01):
    nu = rng.normal(0, 2.0, (K, P))
    Phi = np.stack([(M - m) * 0.1 * rng.uniform(0, 1, (K, P)) * rng.choice([-1, 1], (K, P)) for m in range(M)], axis=2) \
        if M > 0 else np.zeros((K, P, 0))
    eta = 0.1 * rng.normal(0, 1, (P, D, K)) if D else None
    xi = 0.1 * rng.normal(0, 1, (K, P, D, M)) if D else None
    return dict(nu=np.asfortranarray(nu), Phi=np.asfortranarray(Phi), eta=eta, xi=xi, sigma_sq=sigma_sq)


def make_state(rng, n, K, M, conc=10.0):
    pi = rng.dirichlet(np.ones(K))
    Z = rng.dirichlet(conc * pi, size=n)
    chi = rng.normal(0, 1, (n, M))
    return pi, np.asfortranarray(Z), np.asfortranarray(chi)


def theta(par, Z, chi, X=None):
    """theta_i = sum_k Z_ik (nu_k + eta_k x_i + sum_m chi_im (phi_km + xi_km x_i)); n x P"""
    n, K = Z.shape
    P = par["nu"].shape[1]
    M = chi.shape[1]
    th = np.zeros((n, P))
    for k in range(K):
        a = np.tile(par["nu"][k], (n, 1))
        if X is not None:
            a = a + X @ par["eta"][:, :, k].T
        for m in range(M):
            f = np.tile(par["Phi"][k, :, m], (n, 1))
            if X is not None:
                f = f + X @ par["xi"][k][:, :, m].T
            a = a + chi[:, [m]] * f
        th += Z[:, [k]] * a
    return th


def functional_common(seed, n, T, K, P, M, D=0, sigma_sq=0.01, degree=3):
    """Common-grid functional data set: returns dict with y (n x T), t, B (T x P), params, state."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 1000.0, T)
    ik = equispaced_internal(P, degree)
    B = bspline_design(t, ik, degree)
    par = make_params(rng, K, P, M, D, sigma_sq)
    pi, Z, chi = make_state(rng, n, K, M)
    X = np.asfortranarray(rng.normal(0, 1, (n, D))) if D else None
    th = theta(par, Z, chi, X)
    y = th @ B.T + rng.normal(0, np.sqrt(sigma_sq), (n, T))
    return dict(n=n, T=T, K=K, P=P, M=M, D=D, t=t, internal_knots=ik, degree=degree, boundary=(0.0, 1000.0),
                B=B, y=y, X=X, par=par, pi=pi, Z=Z, chi=chi)


def functional_ragged(seed, n, K, P, M, D=0, sigma_sq=0.01, lo=30, hi=50, degree=3):
    rng = np.random.default_rng(seed)
    ik = equispaced_internal(P, degree)
    ni = rng.integers(lo, hi + 1, n)
    off = np.concatenate([[0], np.cumsum(ni)]).astype(np.int64)
    t = np.concatenate([np.sort(rng.uniform(0, 1000.0, m)) for m in ni])
    B = bspline_design(t, ik, degree)
    par = make_params(rng, K, P, M, D, sigma_sq)
    pi, Z, chi = make_state(rng, n, K, M)
    X = np.asfortranarray(rng.normal(0, 1, (n, D))) if D else None
    th = theta(par, Z, chi, X)
    y = np.concatenate([B[off[i]:off[i + 1]] @ th[i] for i in range(n)]) + rng.normal(0, np.sqrt(sigma_sq), off[-1])
    return dict(n=n, K=K, P=P, M=M, D=D, t=t, off=off, internal_knots=ik, degree=degree, boundary=(0.0, 1000.0),
                B=B, y=y, X=X, par=par, pi=pi, Z=Z, chi=chi)


def multivariate(seed, n, R, K, M, D=0, sigma_sq=0.01):
    rng = np.random.default_rng(seed)
    par = make_params(rng, K, R, M, D, sigma_sq)
    pi, Z, chi = make_state(rng, n, K, M)
    X = np.asfortranarray(rng.normal(0, 1, (n, D))) if D else None
    th = theta(par, Z, chi, X)
    y = np.asfortranarray(th + rng.normal(0, np.sqrt(sigma_sq), (n, R)))
    return dict(n=n, K=K, P=R, M=M, D=D, y=y, X=X, par=par, pi=pi, Z=Z, chi=chi)


def tensor_design(t2, internal_per_dim, degree=3, boundary=(0.0, 1000.0)):
    """Row-wise tensor product of two clamped B-spline bases (reference TensorBSpline, BSplines.h:18-62):
    column index = i0 * P1 + i1 (dimension 0 slowest)."""
    B0 = bspline_design(t2[:, 0], internal_per_dim[0], degree, boundary)
    B1 = bspline_design(t2[:, 1], internal_per_dim[1], degree, boundary)
    return (B0[:, :, None] * B1[:, None, :]).reshape(t2.shape[0], -1)


def hd_common(seed, n, K, M, side=32, p_side=20, sigma_sq=0.01, degree=3):
    """High-dimensional functional data (BHDFMMM, BASELINE config 5): a side x side grid on [0, 1000]^2 and
    a p_side x p_side tensor-product cubic basis (P = p_side^2), shared by all functions."""
    rng = np.random.default_rng(seed)
    g = np.linspace(0.0, 1000.0, side)
    t2 = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
    ik = equispaced_internal(p_side, degree)
    B = tensor_design(t2, (ik, ik), degree)
    P, T = p_side * p_side, side * side
    par = make_params(rng, K, P, M, 0, sigma_sq)
    pi, Z, chi = make_state(rng, n, K, M)
    th = theta(par, Z, chi)
    y = th @ B.T + rng.normal(0, np.sqrt(sigma_sq), (n, T))
    return dict(n=n, T=T, K=K, P=P, M=M, D=0, t=t2, internal_knots=ik, degree=degree, boundary=(0.0, 1000.0),
                B=B, y=y, X=None, par=par, pi=pi, Z=Z, chi=chi, p_side=p_side, side=side)
