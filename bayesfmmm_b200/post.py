"""Post-processing over the reference's stored-sample files, on the device.

`conditional_predictive_ordinates` mirrors ConditionalPredictiveOrdinates (src/PostProcessing.cpp:6330-6516):
it reads the batches `Nu{q}.txt, Phi{q}.txt, Z{q}.txt, Sigma{q}.txt` (and `Eta{q}.txt, Xi{q}.txt` for the
covariate-adjusted model) the sampler wrote -- the same files, in the same layout, that
`BFMMM_warm_start` writes (BFMMM.h:1680-1746) -- drops the first floor(burnin_prop * n) stored iterations and
accumulates calcLikelihoodCPO (CalculateLikelihood.h:344-385) on the engine that holds the data."""
from __future__ import annotations

import math
import os

import numpy as np

from . import io as bio


def conditional_predictive_ordinates(engine, directory, n_files, burnin_prop=0.1, cov_adj=False, log_cpo=True):
    if n_files <= 0:
        raise ValueError("'n_files' must be greater than 0")
    if not (0 <= burnin_prop < 1):
        raise ValueError("'burnin_prop' must be between 0 and 1")
    ld = lambda name, q: bio.load(os.path.join(directory, f"{name}{q}.txt"))   # noqa: E731
    nu, Phi, Z, chi, sig, eta, xi = [], [], [], [], [], [], []
    for q in range(n_files):
        nu_q, Z_q, chi_q = ld("Nu", q), ld("Z", q), ld("Chi", q)                # cubes K x P x S, n x K x S, n x M x S
        Phi_q = ld("Phi", q)                                                    # field (S, 1) of K x P x M cubes
        s_q = ld("Sigma", q).ravel()
        S = s_q.size
        nu += [nu_q[:, :, l] for l in range(S)]
        Z += [Z_q[:, :, l] for l in range(S)]
        chi += [chi_q[:, :, l] for l in range(S)]
        Phi += [Phi_q[l, 0] for l in range(S)]
        sig += list(s_q)
        if cov_adj:
            eta_q, xi_q = ld("Eta", q), ld("Xi", q)                             # fields (S, 1) of P x D x K, (S, K) of P x D x M
            eta += [eta_q[l, 0] for l in range(S)]
            xi += [np.stack([xi_q[l, k] for k in range(xi_q.shape[1])]) for l in range(S)]
    n_iter = len(sig)
    first = int(math.floor(burnin_prop * n_iter))
    engine.cpo_reset()
    for l in range(first, n_iter):
        engine.set_state(Z[l], chi[l])
        if cov_adj:
            engine.set_globals(nu[l], Phi[l], float(sig[l]), eta=eta[l], xi=xi[l])
        else:
            engine.set_globals(nu[l], Phi[l], float(sig[l]))
        engine.cpo_accumulate()
    return engine.cpo_get(log_scale=log_cpo)


# ------------------------------------------------------------------------------------------ credible intervals
def quantiles(draws, probs, device=0):
    """Quantiles over axis 0 of `draws` (S x ...), Armadillo's definition, sorted on the device
    (include/bfmmm_post.h: bfmmm_quantiles).  Returns an array of shape (len(probs), ...)."""
    import ctypes as C
    from ._lib import dp, load_library
    from .engine import EngineError
    x = np.ascontiguousarray(np.asarray(draws, dtype=np.float64))
    S = x.shape[0]
    R = int(np.prod(x.shape[1:])) if x.ndim > 1 else 1
    p = np.ascontiguousarray(np.asarray(probs, dtype=np.float64).ravel())
    out = np.empty((p.size, R))
    lib = load_library()
    rc = lib.bfmmm_quantiles(x.ctypes.data_as(dp), C.c_int64(S), C.c_int64(R), p.ctypes.data_as(dp), C.c_int(p.size),
                             out.ctypes.data_as(dp), C.c_int(device))
    if rc != 0:
        raise EngineError(lib.bfmmm_last_error().decode())
    return out.reshape((p.size,) + x.shape[1:])


def _check_common(n_files, alpha, burnin_prop):
    if n_files <= 0:
        raise ValueError("'n_files' must be greater than 0")
    if not (0 <= alpha < 1):
        raise ValueError("'alpha' must be between 0 and 1")
    if not (0 <= burnin_prop < 1):
        raise ValueError("'burnin_prop' must be between 0 and 1")


def _stack(directory, name, n_files):
    """All stored draws of one parameter, draw index first (the reference concatenates the batches in order)."""
    parts = []
    for q in range(n_files):
        a = bio.load(os.path.join(directory, f"{name}{q}.txt"))
        if a.ndim == 5:                      # field (S, fc) of cubes
            parts.append(a)
        elif a.ndim == 3:                    # cube r x c x S
            parts.append(np.moveaxis(a, 2, 0))
        else:                                # S x 1 (vec) or S x c
            parts.append(a)
    return np.concatenate(parts, axis=0)


def _keep(x, burnin_prop):
    """The last round(S (1 - burnin_prop)) draws (e.g. src/PostProcessing.cpp:3462-3463)."""
    S = x.shape[0]
    m = int(round(S * (1 - burnin_prop)))
    return x[S - m:]


def _rescale_mats(Z):
    """The reference's rescaling (K = 2 only, e.g. src/PostProcessing.cpp:3528-3541): row i of the transform is the
    membership row of the function with the largest Z_.i in that draw."""
    S, n, K = Z.shape
    idx = Z.argmax(axis=1)                                   # S x K
    return np.stack([Z[s, idx[s], :] for s in range(S)])     # S x K x K


def sigma_ci(directory, n_files, alpha=0.05, burnin_prop=0.1, device=0):
    """SigmaCI (src/PostProcessing.cpp:3435-3480).  Note the reference returns the MEDIAN as `CI_Lower` (:3472);
    this function returns the alpha/2 quantile there and the reference's value as `CI_Lower_reference`."""
    _check_common(n_files, alpha, burnin_prop)
    sig = _keep(_stack(directory, "Sigma", n_files).reshape(-1, 1), burnin_prop)
    q = quantiles(sig, [alpha / 2, 0.5, 1 - alpha / 2], device)[:, 0]
    return {"CI_Upper": q[2], "CI_50": q[1], "CI_Lower": q[0], "CI_Lower_reference": q[1]}


def z_ci(directory, n_files, alpha=0.05, rescale=True, burnin_prop=0.1, device=0):
    """ZCI (src/PostProcessing.cpp:3505-3592): the n x K quantiles run one thread block per element on the device."""
    _check_common(n_files, alpha, burnin_prop)
    Z = _stack(directory, "Z", n_files)                      # S x n x K
    if rescale and Z.shape[2] > 2:
        rescale = False                                      # "Rescale property cannot be used for K > 2" (:3518-3523)
    if rescale:
        T = _rescale_mats(Z)
        Z = np.stack([np.linalg.solve(T[s].T, Z[s].T).T for s in range(Z.shape[0])])    # :3540
    q = quantiles(_keep(Z, burnin_prop), [alpha / 2, 0.5, 1 - alpha / 2], device)
    return {"CI_Upper": q[2], "CI_50": q[1], "CI_Lower": q[0]}


def _band(f_samp, alpha, simultaneous, device):
    """pointwise quantile band, or the reference's simultaneous band  mean +- q_{1-alpha}(max_j |f - mean| / sd) sd"""
    if not simultaneous:
        q = quantiles(f_samp, [alpha / 2, 0.5, 1 - alpha / 2], device)
        return {"CI_Upper": q[2], "CI_50": q[1], "CI_Lower": q[0]}
    mean, sd = f_samp.mean(axis=0), f_samp.std(axis=0, ddof=1)
    C = np.max(np.abs((f_samp - mean) / sd).reshape(f_samp.shape[0], -1), axis=1)
    qc = quantiles(C.reshape(-1, 1), [1 - alpha], device)[0, 0]
    return {"CI_Upper": mean + qc * sd, "CI_50": mean, "CI_Lower": mean - qc * sd}


def f_mean_ci(directory, n_files, time, basis_degree, boundary_knots, internal_knots, k, alpha=0.05, rescale=True,
              simultaneous=False, burnin_prop=0.1, trans_mats=None, device=0):
    """FMeanCI without covariates (src/PostProcessing.cpp:99-480): credible band of the k-th mean function B(t) nu_k."""
    from . import basis as bfbasis
    _check_common(n_files, alpha, burnin_prop)
    if basis_degree < 1:
        raise ValueError("'basis_degree' must be an integer greater than or equal to 1")
    nu = _keep(_stack(directory, "Nu", n_files), burnin_prop)            # S x K x P
    K = nu.shape[1]
    if k <= 0 or k > K:
        raise ValueError("'k' must be positive and at most the number of clusters in the model")
    if rescale and K > 2:
        rescale = False
    if rescale:
        T = _rescale_mats(_keep(_stack(directory, "Z", n_files), burnin_prop))
        nu = np.einsum("sij,sjp->sip", T, nu)                             # nu <- transform_mat nu (:221)
    elif trans_mats is not None:
        tm = np.asarray(trans_mats, dtype=np.float64).reshape(nu.shape[0], K, K)
        nu = np.einsum("sij,sjp->sip", tm, nu)
    B = bfbasis.bspline_basis(np.asarray(time, float), np.asarray(internal_knots, float), basis_degree, tuple(boundary_knots))
    f_samp = nu[:, k - 1, :] @ B.T                                        # S x T
    return _band(f_samp, alpha, simultaneous, device)


def f_cov_ci(directory, n_files, time1, time2, basis_degree, boundary_knots, internal_knots, l, m, alpha=0.05,
             rescale=True, simultaneous=False, burnin_prop=0.1, trans_mats=None, device=0):
    """FCovCI without covariates (src/PostProcessing.cpp:1781-2300): credible band of the covariance surface between
    clusters l and m,  C(t1, t2) = sum_j (B(t1) phi_lj)(B(t2) phi_mj)."""
    from . import basis as bfbasis
    _check_common(n_files, alpha, burnin_prop)
    Phi = _keep(_stack(directory, "Phi", n_files)[:, 0], burnin_prop)     # S x K x P x M
    K = Phi.shape[1]
    if min(l, m) <= 0 or max(l, m) > K:
        raise ValueError("'l' and 'm' must be positive and at most the number of clusters in the model")
    if rescale and K > 2:
        rescale = False
    if rescale:
        T = _rescale_mats(_keep(_stack(directory, "Z", n_files), burnin_prop))
        Phi = np.einsum("sij,sjpm->sipm", T, Phi)
    elif trans_mats is not None:
        tm = np.asarray(trans_mats, dtype=np.float64).reshape(Phi.shape[0], K, K)
        Phi = np.einsum("sij,sjpm->sipm", tm, Phi)
    kn = np.asarray(internal_knots, float)
    B1 = bfbasis.bspline_basis(np.asarray(time1, float), kn, basis_degree, tuple(boundary_knots))
    B2 = bfbasis.bspline_basis(np.asarray(time2, float), kn, basis_degree, tuple(boundary_knots))
    a = np.einsum("tp,spj->stj", B1, Phi[:, l - 1])                      # S x T1 x M
    b = np.einsum("tp,spj->stj", B2, Phi[:, m - 1])
    cov = np.einsum("saj,sbj->sab", a, b)                                 # S x T1 x T2
    return _band(cov, alpha, simultaneous, device)
