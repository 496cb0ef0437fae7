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
