"""Edge cases and full-size properties of the C ABI on the GPU.

Small/odd shapes are compared with the oracle; at BASELINE.json's full size (n = 1e6 functions,
K=3, P=20, M=3, T=200) the oracle would take hours, so size-independent properties are checked:
additivity of every reduced statistic over shards, agreement of the fused reductions with
recomputations from the returned state, row sums of Z, reproducibility."""
import numpy as np
import pytest

import bayesfmmm_b200 as bf
from bayesfmmm_b200.engine import FUNCTIONAL, MULTIVARIATE
from oracle import oracle as orc
from tests import synth
from tests.gpu_util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _mk(s, **kw):
    eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], B=s["B"], T=s["T"], X=s["X"], **kw)
    eng.set_state(s["Z"], s["chi"])
    p = s["par"]
    eng.set_globals(p["nu"], p["Phi"], p["sigma_sq"], eta=p["eta"], xi=p["xi"])
    return eng


def _oracle(s):
    n, T = s["n"], s["T"]
    d = orc.Data(n=n, K=s["K"], P=s["P"], M=s["M"], y=s["y"].ravel(), B=np.tile(s["B"], (n, 1)),
                 off=np.arange(n + 1, dtype=np.int64) * T, X=s["X"])
    p = s["par"]
    st = orc.State(nu=p["nu"], Phi=p["Phi"], Z=s["Z"], chi=s["chi"], sigma_sq=p["sigma_sq"], eta=p["eta"], xi=p["xi"])
    return d, st


@pytest.mark.parametrize("n", [1, 2, 7, 9, 129, 257])
def test_odd_function_counts(n):
    """n not a multiple of the 8-function chunks / 128-thread blocks: padding never leaks into results."""
    s = synth.functional_common(seed=40 + n, n=n, T=33, K=3, P=8, M=2)
    d, st = _oracle(s)
    eng = _mk(s)
    rng = np.random.default_rng(n)
    assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
    gam = np.asfortranarray(rng.gamma(5000.0 * s["Z"])); u = rng.uniform(size=n)
    Zo, acc, took = orc.update_z(d, st, s["pi"], 1.1, 5000.0, gam, u)
    slz, nacc = eng.update_z(s["pi"], 1.1, 5000.0, 1.0, gam=gam, u=u)
    Zg, _ = eng.get_state(chi=False)
    assert np.array_equal(Zg, Zo) and nacc == took.sum() and rel(slz, np.log(Zo).sum(axis=0)) < TOL
    eng.set_state(s["Z"], s["chi"])
    eps = np.asfortranarray(rng.normal(size=(n, 2)))
    eng.update_chi(1.0, eps=eps)
    assert rel(eng.get_state(Z=False)[1], orc.update_chi(d, st, eps)) < TOL
    W, R = eng.suffstats()
    assert np.allclose(W, W.T) and np.all(np.isfinite(R))
    eng.close()


@pytest.mark.parametrize("K,M,P,D", [(6, 6, 12, 0), (2, 1, 5, 0), (4, 5, 9, 4), (5, 2, 64, 1), (3, 3, 130, 0)])
def test_largest_and_smallest_supported_shapes(K, M, P, D):
    s = synth.functional_common(seed=K * 10 + M, n=50, T=max(2 * P, 40), K=K, P=P, M=M, D=D)
    d, st = _oracle(s)
    eng = _mk(s)
    rng = np.random.default_rng(1)
    assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
    eps = np.asfortranarray(rng.normal(size=(50, M)))
    ssr_after = eng.update_chi(1.0, eps=eps)
    chi_o = orc.update_chi(d, st, eps)
    assert rel(eng.get_state(Z=False)[1], chi_o) < TOL      # measured 2e-15 .. 3e-15 on these shapes (B200)
    eng.set_state(s["Z"], s["chi"])
    gam = np.asfortranarray(rng.gamma(20000.0 * s["Z"])); u = rng.uniform(size=50)
    Zo, acc, _ = orc.update_z(d, st, s["pi"], 1.0, 20000.0, gam, u)
    eng.update_z(s["pi"], 1.0, 20000.0, 1.0, gam=gam, u=u)
    ok = ~(np.abs(np.log(u) - acc) <= 1e-7)
    assert np.array_equal(eng.get_state(chi=False)[0][ok], Zo[ok])
    eng.close()


def test_invalid_configurations_fail_loudly():
    s = synth.functional_common(seed=1, n=10, T=20, K=3, P=6, M=2)
    with pytest.raises(bf.EngineError):      # K outside the instantiated range
        bf.Engine(model=FUNCTIONAL, n=10, K=7, P=6, M=2, y=s["y"], B=s["B"], T=20)
    with pytest.raises(bf.EngineError):      # no basis and no spline description
        bf.Engine(model=FUNCTIONAL, n=10, K=3, P=6, M=2, y=s["y"], T=20)
    with pytest.raises(bf.EngineError):      # bad device
        bf.Engine(model=FUNCTIONAL, n=10, K=3, P=6, M=2, y=s["y"], B=s["B"], T=20, device=99)
    eng = bf.Engine(model=FUNCTIONAL, n=10, K=3, P=6, M=2, y=s["y"], B=s["B"], T=20)
    with pytest.raises(bf.EngineError):      # sigma^2 must be positive
        eng.set_globals(s["par"]["nu"], s["par"]["Phi"], 0.0)
    with pytest.raises(bf.EngineError):      # gam without u
        eng._chk(eng._lib.bfmmm_update_z(eng._h, np.ones(3).ctypes.data_as(bf._lib.dp), bf._lib.C.c_double(1.0),
                                         bf._lib.C.c_double(10.0), bf._lib.C.c_double(1.0),
                                         np.ones((10, 3)).ctypes.data_as(bf._lib.dp), None, None, None))
    eng.close()


def test_zero_memberships_follow_the_reference_rule():
    """A current Z_ik <= 0 forces acceptance (UpdateMixedMembership.h:170-174) and a non-positive
    Dirichlet concentration is replaced by 10 (Distributions.h:24-28)."""
    s = synth.functional_common(seed=3, n=16, T=30, K=3, P=6, M=2)
    s["Z"][:, 0] = 0.0
    s["Z"][:, 1:] /= s["Z"][:, 1:].sum(axis=1, keepdims=True)
    d, st = _oracle(s)
    eng = _mk(s)
    rng = np.random.default_rng(0)
    gam = np.asfortranarray(rng.gamma(np.where(s["Z"] > 0, 100.0 * s["Z"], 10.0))); u = rng.uniform(size=16)
    Zo, acc, took = orc.update_z(d, st, s["pi"], 1.0, 100.0, gam, u)
    slz, nacc = eng.update_z(s["pi"], 1.0, 100.0, 1.0, gam=gam, u=u)
    assert took.all() and nacc == 16
    assert np.array_equal(eng.get_state(chi=False)[0], Zo)
    # device RNG path with zero memberships also proposes from Gamma(10) and accepts
    eng.set_state(s["Z"], s["chi"])
    eng.seed(5, 1)
    _, nacc = eng.update_z(s["pi"], 1.0, 100.0, 1.0)
    Zg = eng.get_state(chi=False)[0]
    assert nacc == 16 and np.all(Zg > 0) and np.allclose(Zg.sum(axis=1), 1.0)
    eng.close()


def test_full_size_properties():
    """n = 1e6, K=3, P=20, M=3, T=200 (the metric's shape; generated in 4 chunks of 250k functions)."""
    n, T, K, P, M = 1_000_000, 200, 3, 20, 3
    rng = np.random.default_rng(0)
    t = np.linspace(0, 1000.0, T); ik = synth.equispaced_internal(P, 3); B = synth.bspline_design(t, ik, 3)
    par = synth.make_params(np.random.default_rng(1), K, P, M, 0, 0.01)
    pi = np.array([0.2, 0.3, 0.5])
    Z = np.asfortranarray(rng.dirichlet(10 * pi, size=n)); chi = np.asfortranarray(rng.normal(size=(n, M)))
    y = np.empty((n, T))
    for c in range(4):
        sl = slice(c * n // 4, (c + 1) * n // 4)
        y[sl] = synth.theta(par, Z[sl], chi[sl]) @ B.T + 0.1 * rng.standard_normal((n // 4, T))
    full = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=y, T=T, t=t, degree=3, internal_knots=ik, boundary=(0.0, 1000.0))
    full.set_state(Z, chi); full.set_globals(par["nu"], par["Phi"], 0.01)
    ssr_full = full.ssr()[0]
    W_full, R_full = full.suffstats()
    # (1) additivity over shards: the same data split in two engines
    h = 400_003
    parts = []
    for lo, hi in ((0, h), (h, n)):
        e = bf.Engine(model=FUNCTIONAL, n=hi - lo, K=K, P=P, M=M, y=y[lo:hi], B=B, T=T, global_offset=lo)
        e.set_state(Z[lo:hi], chi[lo:hi]); e.set_globals(par["nu"], par["Phi"], 0.01)
        parts.append((e.ssr()[0],) + e.suffstats())
        e.close()
    assert rel(parts[0][0] + parts[1][0], ssr_full) < 1e-11
    assert rel(parts[0][1] + parts[1][1], W_full) < 1e-11 and rel(parts[0][2] + parts[1][2], R_full) < 1e-11
    # (2) the statistics are the plain matrix products
    Wm = np.concatenate([np.stack([Z[:, k]] + [Z[:, k] * chi[:, m] for m in range(M)], axis=1) for k in range(K)], axis=1)
    assert rel(W_full, Wm.T @ Wm) < 1e-11
    assert rel(R_full, (y @ B).T @ Wm) < 1e-10
    # (3) a residual check that does not need the oracle: SSR from the definition on a 2000-function sample
    idx = rng.choice(n, 2000, replace=False)
    sub = bf.Engine(model=FUNCTIONAL, n=2000, K=K, P=P, M=M, y=y[idx], B=B, T=T)
    sub.set_state(Z[idx], chi[idx]); sub.set_globals(par["nu"], par["Phi"], 0.01)
    direct = ((y[idx] - synth.theta(par, Z[idx], chi[idx]) @ B.T) ** 2).sum()
    assert rel(sub.ssr()[0], direct) < 1e-11
    sub.close()
    # (4) the Z step: rows stay on the simplex, the fused sum of log Z equals the recomputed one, and
    # the same seed reproduces the same state; the chi step's fused SSR equals a fresh SSR pass
    full.seed(7, 0)
    slz, nacc = full.update_z(pi, 1.0, 10000.0)
    Z1, _ = full.get_state(chi=False)
    assert np.max(np.abs(Z1.sum(axis=1) - 1)) < 1e-12 and Z1.min() > 0
    assert rel(slz, np.log(Z1).sum(axis=0)) < 1e-11
    assert 0.2 < nacc / n < 0.95
    changed = np.any(Z1 != Z, axis=1).sum()
    assert changed == nacc
    full.set_state(Z, None); full.seed(7, 0)
    full.update_z(pi, 1.0, 10000.0)
    assert np.array_equal(full.get_state(chi=False)[0], Z1)
    ssr_after = full.update_chi(1.0)
    assert rel(ssr_after, full.ssr()[0]) < 1e-11
    full.close()


@pytest.mark.parametrize("kind", ["no_support", "fewer_points_than_basis"])
def test_rank_deficient_basis(kind):
    """A basis whose Gram matrix is singular on the grid (a basis function without support, or T < P):
    the reference works with B directly, the engine falls back to an eigen-whitening of rank r < P."""
    rng = np.random.default_rng(7)
    n, K, P, M = 40, 3, 8, 2
    ik = synth.equispaced_internal(P, 3)
    if kind == "no_support":
        t = np.linspace(0.0, 450.0, 30)            # the last basis functions vanish on [0, 450]
    else:
        t = np.linspace(0.0, 1000.0, 5)            # T = 5 < P = 8
    B = synth.bspline_design(t, ik, 3)
    assert np.linalg.matrix_rank(B) < P
    par = synth.make_params(rng, K, P, M, 0, 0.01)
    pi, Z, chi = synth.make_state(rng, n, K, M)
    y = synth.theta(par, Z, chi) @ B.T + 0.1 * rng.normal(size=(n, len(t)))
    s = dict(n=n, K=K, P=P, M=M, T=len(t), B=B, y=y, X=None, par=par, Z=Z, chi=chi, pi=pi)
    d, st = _oracle(s)
    eng = _mk(s)
    assert rel(eng.ssr()[0], orc.ssr(d, st)[0]) < TOL
    eps = np.asfortranarray(rng.normal(size=(n, M)))
    ssr_after = eng.update_chi(1.0, eps=eps)
    chi_o = orc.update_chi(d, st, eps)
    # rank-deficient Gram: the cache is whitened by eigenvectors with a relative eigenvalue cut-off, the oracle
    # works per observed point; 1e-9 is the bar here (not re-measured this round)
    assert rel(eng.get_state(Z=False)[1], chi_o) < 1e-9
    eng.set_state(Z, chi)
    gam = np.asfortranarray(rng.gamma(20000.0 * Z)); u = rng.uniform(size=n)
    Zo, acc, _ = orc.update_z(d, st, pi, 1.0, 20000.0, gam, u)
    eng.update_z(pi, 1.0, 20000.0, 1.0, gam=gam, u=u)
    ok = ~(np.abs(np.log(u) - acc) <= 1e-7)
    assert np.array_equal(eng.get_state(chi=False)[0][ok], Zo[ok])
    eng.set_state(Z, chi)
    W, R = eng.suffstats()
    Wm = np.concatenate([np.stack([Z[:, k]] + [Z[:, k] * chi[:, m] for m in range(M)], axis=1) for k in range(K)], axis=1)
    assert rel(R, (y @ B).T @ Wm) < 1e-9 and rel(eng.gram(), B.T @ B) < 1e-12
    eng.close()


def test_overlapped_state_readback():
    """bfmmm_get_state_begin / _wait return the state as of the call, even when updates are queued behind it."""
    s = synth.functional_common(seed=77, n=1000, T=40, K=3, P=10, M=2)
    eng = _mk(s)
    n = s["n"]
    Z1, c1 = np.zeros((n, 3), order="F"), np.zeros((n, 2), order="F")
    Z2, c2 = np.zeros((n, 3), order="F"), np.zeros((n, 2), order="F")
    eng.get_state_begin(Z1, c1)
    eng.update_chi_async()                       # queued behind the snapshot: must not leak into slice 1
    eng.update_z_async(s["pi"], 1.0, 1000.0)
    eng.get_state_begin(Z2, c2)
    eng.get_state_wait()
    assert np.array_equal(Z1, s["Z"]) and np.array_equal(c1, s["chi"])
    Zn, cn = eng.get_state()
    assert np.array_equal(Z2, Zn) and np.array_equal(c2, cn)
    assert not np.array_equal(c2, c1)
    eng.get_state_wait()                         # idempotent
    eng.close()
