"""Post-processing over stored draws: the device quantile routine against NumPy's type-5 (Hazen)
quantile -- the definition Armadillo's quantile() documents -- and the four credible-interval functions on the
reference's own stored chains (tests/golden/Functional_trace = inst/test-data/Functional_trace)."""
import os

import numpy as np
import pytest

from bayesfmmm_b200 import io as bio
from bayesfmmm_b200 import post

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TRACE = os.path.join(GOLD, "Functional_trace") + os.sep


@pytest.mark.parametrize("S,R", [(1, 5), (2, 3), (75, 80), (150, 1), (1000, 257), (4097, 9)])
def test_device_quantiles_match_hazen(S, R):
    rng = np.random.default_rng(S * 31 + R)
    x = rng.normal(size=(S, R))
    x[:, 0] = np.round(x[:, 0], 1)                       # ties
    p = np.array([0.0, 0.025, 0.05, 0.5, 0.9, 0.975, 1.0, 1e-6])
    got = post.quantiles(x, p)
    ref = np.quantile(x, p, axis=0, method="hazen")
    assert np.max(np.abs(got - ref)) <= 1e-15 * (1 + np.max(np.abs(ref)))


def test_sigma_ci_and_z_ci_on_the_reference_chain():
    ci = post.sigma_ci(TRACE, 1)
    sig = bio.load(TRACE + "Sigma0.txt").ravel()
    keep = sig[len(sig) - int(round(len(sig) * 0.9)):]
    q = np.quantile(keep, [0.025, 0.5, 0.975], method="hazen")
    assert abs(ci["CI_Lower"] - q[0]) < 1e-18 and abs(ci["CI_50"] - q[1]) < 1e-18 and abs(ci["CI_Upper"] - q[2]) < 1e-18
    assert ci["CI_Lower_reference"] == ci["CI_50"]                        # the reference's SigmaCI returns q(1) twice
    Z = np.moveaxis(bio.load(TRACE + "Z0.txt"), 2, 0)                     # S x n x K
    z = post.z_ci(TRACE, 1, rescale=False)
    keepz = Z[Z.shape[0] - int(round(Z.shape[0] * 0.9)):]
    qz = np.quantile(keepz, [0.025, 0.5, 0.975], axis=0, method="hazen")
    assert np.array_equal(z["CI_Lower"], qz[0]) or np.max(np.abs(z["CI_Lower"] - qz[0])) < 1e-15
    assert np.max(np.abs(z["CI_50"] - qz[1])) < 1e-15 and np.max(np.abs(z["CI_Upper"] - qz[2])) < 1e-15
    # rescaled (K = 2): in every draw some function is entirely in each cluster
    zr = post.z_ci(TRACE, 1, rescale=True)
    assert zr["CI_50"].shape == (40, 2) and np.all(zr["CI_Lower"] <= zr["CI_Upper"])
    assert abs(zr["CI_50"].max() - 1.0) < 1e-9 and np.allclose(zr["CI_50"].sum(axis=1), 1.0, atol=0.05)


def test_function_bands_on_the_reference_chain():
    t = np.arange(0.0, 1000.0, 10.0)
    kw = dict(basis_degree=3, boundary_knots=(0.0, 1000.0), internal_knots=[250.0, 500.0, 750.0])
    m = post.f_mean_ci(TRACE, 1, t, k=1, rescale=False, **kw)
    nu = np.moveaxis(bio.load(TRACE + "Nu0.txt"), 2, 0)
    from tests import synth
    B = synth.bspline_design(t, [250.0, 500.0, 750.0], 3)
    f = nu[nu.shape[0] - int(round(nu.shape[0] * 0.9)):, 0, :] @ B.T
    q = np.quantile(f, [0.025, 0.5, 0.975], axis=0, method="hazen")
    assert np.max(np.abs(m["CI_50"] - q[1])) < 1e-12 and np.max(np.abs(m["CI_Upper"] - q[2])) < 1e-12
    ms = post.f_mean_ci(TRACE, 1, t, k=2, rescale=True, simultaneous=True, **kw)
    assert np.all(ms["CI_Lower"] < ms["CI_50"]) and np.all(ms["CI_50"] < ms["CI_Upper"])
    c = post.f_cov_ci(TRACE, 1, t[::5], t[::10], l=1, m=2, rescale=False, **kw)
    assert c["CI_50"].shape == (20, 10) and np.all(c["CI_Lower"] <= c["CI_Upper"])
    cd = post.f_cov_ci(TRACE, 1, t[::5], t[::5], l=1, m=1, rescale=False, **kw)
    assert np.all(np.diag(cd["CI_Lower"]) >= 0)                           # variances
