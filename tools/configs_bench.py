"""Sweep timings of the other BASELINE.json configurations (they are parity-test cases, not bench lines;
this script only records how the same engine behaves on their shapes).  usage: configs_bench.py [c1 c2 c3 c4 c5]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesfmmm_b200 as bf                                  # noqa: E402
from bayesfmmm_b200 import basis as bfbasis                  # noqa: E402
from bayesfmmm_b200.engine import FUNCTIONAL, MULTIVARIATE   # noqa: E402
from tests import synth                                      # noqa: E402


def run(name, eng, smp, sweep, steps=30, warm=5):
    for _ in range(warm):
        smp.step(sweep)
    eng.sync()
    p0 = smp.profile()
    t0 = time.perf_counter()
    for _ in range(steps):
        smp.step(sweep)
    eng.sync()
    dt = (time.perf_counter() - t0) / steps
    p1 = smp.profile()
    g = smp.get()
    out = {"config": name, "ms_per_sweep": dt * 1e3, "sweeps_per_s": 1 / dt, "sigma_sq": g["sigma_sq"], "loglik": g["loglik"],
           "host_ms": (p1["host_s"] - p0["host_s"]) / steps * 1e3, "wait_ms": (p1["wait_s"] - p0["wait_s"]) / steps * 1e3}
    print(json.dumps(out), flush=True)
    smp.close(); eng.close()


def common(name, n, T, K, P, M, seed):
    s = synth.functional_common(seed=seed, n=n, T=T, K=K, P=P, M=M)
    eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=s["y"], T=T, t=s["t"], degree=3,
                    internal_knots=s["internal_knots"], boundary=(0.0, 1000.0))
    eng.set_state(s["Z"], s["chi"])
    smp = bf.Sampler(eng, hyper=bf.default_hyper(True), n_total=n, Pmat=bfbasis.pmat_rw1(P), seed=1)
    smp.set(nu=s["par"]["nu"], Phi=s["par"]["Phi"], sigma_sq=0.01, pi=s["pi"], alpha3=1.0)
    run(name, eng, smp, bf.SWEEP_FULL)


def c3(n=1_000_000):
    s = synth.multivariate(seed=3, n=n, R=64, K=3, M=4)
    eng = bf.Engine(model=MULTIVARIATE, n=n, K=3, P=64, M=4, y=s["y"])
    eng.set_state(s["Z"], s["chi"])
    smp = bf.Sampler(eng, hyper=bf.default_hyper(False), n_total=n, Pmat=None, seed=1)
    smp.set(nu=s["par"]["nu"], Phi=s["par"]["Phi"], sigma_sq=0.01, pi=s["pi"], alpha3=1.0)
    run(f"C3 multivariate K=3 R=64 M=4 n={n}", eng, smp, bf.SWEEP_FULL)


def c4(n=100_000):
    """covariate-adjusted, ragged grids (n_i ~ U{150..250}); the basis is evaluated on the device from t"""
    from scipy.interpolate import BSpline
    rng = np.random.default_rng(4)
    K, P, M, D = 3, 20, 3, 2
    ik = synth.equispaced_internal(P)
    ni = rng.integers(150, 251, n)
    off = np.concatenate([[0], np.cumsum(ni)]).astype(np.int64)
    N = int(off[-1])
    t = rng.uniform(0, 1000.0, N)
    # sort within each function
    order = np.lexsort((t, np.repeat(np.arange(n), ni)))
    t = t[order]
    par = synth.make_params(rng, K, P, M, D)
    pi, Z, chi = synth.make_state(rng, n, K, M)
    X = np.asfortranarray(rng.normal(0, 1, (n, D)))
    th = synth.theta(par, Z, chi, X)
    Bsp = BSpline.design_matrix(t, synth.clamped_knots(ik, 3, (0.0, 1000.0)), 3, extrapolate=False).tocsr()
    y = np.asarray(Bsp.multiply(np.repeat(th, ni, axis=0)).sum(axis=1)).ravel() + rng.normal(0, 0.1, N)
    eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=y, off=off, t=t, degree=3, internal_knots=ik,
                    boundary=(0.0, 1000.0), X=X, common_grid=False)
    eng.set_state(Z, chi)
    smp = bf.Sampler(eng, hyper=bf.default_hyper(True), n_total=n, Pmat=bfbasis.pmat_rw1(P), seed=1)
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=0.01, pi=pi, alpha3=1.0)
    smp.set_cov(eta=par["eta"], xi=par["xi"])
    run(f"C4 covariate-adjusted D=2 K=3 P=20 M=3 ragged n={n} ({N} points)", eng, smp, bf.SWEEP_FULL, steps=20)


def c5(n=200_000):
    """high-dimensional functional: 20 x 20 tensor-product cubic basis (P = 400) on a 32 x 32 grid"""
    rng = np.random.default_rng(5)
    K, M, P = 4, 3, 400
    g1 = np.linspace(0, 990.0, 32)
    tt = np.stack(np.meshgrid(g1, g1, indexing="ij"), axis=-1).reshape(-1, 2)
    ik = synth.equispaced_internal(20, 3, (0.0, 990.0))
    B = bfbasis.tensor_bspline(tt, [3, 3], [(0.0, 990.0), (0.0, 990.0)], [ik, ik])
    par = synth.make_params(rng, K, P, M)
    pi, Z, chi = synth.make_state(rng, n, K, M)
    th = synth.theta(par, Z, chi)
    y = th @ B.T + rng.normal(0, 0.1, (n, B.shape[0]))
    del th
    eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=y, B=B, T=B.shape[0])
    del y
    eng.set_state(Z, chi)
    Pm = bfbasis.get_P([3, 3], [ik, ik])
    smp = bf.Sampler(eng, hyper=bf.default_hyper(True), n_total=n, Pmat=Pm, seed=1)
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=0.01, pi=pi, alpha3=1.0)
    run(f"C5 high-dimensional K=4 P=400 (20x20 tensor basis) T=1024 n={n}", eng, smp, bf.SWEEP_FULL, steps=10, warm=2)


def cpo(n=1_000_000):
    """CPO accumulation (calcLikelihoodCPO per retained iteration) at the benchmark shape, next to the reference's
    own function (oracle/_ref, dense n_i x n_i covariance per function) on a few functions."""
    s = synth.functional_common(seed=1, n=n, T=200, K=3, P=20, M=3)
    eng = bf.Engine(model=FUNCTIONAL, n=n, K=3, P=20, M=3, y=s["y"], T=200, t=s["t"], degree=3,
                    internal_knots=s["internal_knots"], boundary=(0.0, 1000.0))
    eng.set_state(s["Z"], s["chi"])
    eng.set_globals(s["par"]["nu"], s["par"]["Phi"], 0.01)
    eng.cpo_reset()
    for _ in range(3):
        eng.cpo_accumulate()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(50):
        eng.cpo_accumulate()
    eng.sync()
    dt = (time.perf_counter() - t0) / 50
    out = {"config": f"CPO accumulation, functional K=3 P=20 M=3 n={n} T=200", "ms_per_iteration": dt * 1e3}
    try:
        from oracle import oracle as orc, ref
        if ref.available():
            m = 20
            d = orc.Data(n=m, K=3, P=20, M=3, y=s["y"][:m].ravel(), B=np.tile(s["B"], (m, 1)), off=np.arange(m + 1, dtype=np.int64) * 200)
            st = orc.State(nu=s["par"]["nu"], Phi=s["par"]["Phi"], Z=s["Z"][:m], chi=s["chi"][:m], sigma_sq=0.01)
            t0 = time.perf_counter()
            ref.cpo(d, [st])
            tr = time.perf_counter() - t0
            out["reference_s_per_iteration_extrapolated"] = tr / m * n
            out["reference_sample"] = f"{m} functions, {tr:.3f} s, one host core (oracle/_ref over the shim)"
    except Exception as exc:
        out["reference_sample"] = f"unavailable: {exc}"
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    if "c1" in which:
        common("C1 functional K=2 P=8 M=3 n=40 T=100", 40, 100, 2, 8, 3, 1)
    if "c2" in which:
        common("C2 functional K=3 P=20 M=3 n=100000 T=200", 100_000, 200, 3, 20, 3, 2)
    if "c3" in which:
        c3()
    if "c4" in which:
        c4()
    if "c5" in which:
        c5()
    if "cpo" in which:
        cpo()
