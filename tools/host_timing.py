"""Times the host-side updates of the sampler at a given shape (detached sampler, no GPU work)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bayesfmmm_b200 as bf
from bayesfmmm_b200 import basis as bb
from tests import synth

def run(K, M, nb):
    P = nb * nb
    ik = synth.equispaced_internal(nb, 3, (0.0, 990.0))
    g1 = np.linspace(0, 990.0, 32); tt = np.stack(np.meshgrid(g1, g1, indexing="ij"), axis=-1).reshape(-1, 2)
    B = bb.tensor_bspline(tt, [3, 3], [(0.0, 990.0), (0.0, 990.0)], [ik, ik])
    G = B.T @ B; Pm = bb.get_P([3, 3], [ik, ik])
    n = 2000
    smp = bf.Sampler(hyper=bf.default_hyper(True), n_total=n, Pmat=Pm, dims=(n, K, P, M, 0, 0), G=G, sum_half_total=n * 512.0, n_points_total=n * 1024.0)
    rng = np.random.default_rng(0)
    smp.set(nu=np.asfortranarray(rng.normal(size=(K, P))), Phi=np.asfortranarray(rng.normal(size=(K, P, M)) * 0.1), sigma_sq=0.01, pi=np.ones(K) / K, alpha3=1.0)
    q = K * (M + 1)
    W = rng.normal(size=(n, q)); WtW = W.T @ W; BtYW = rng.normal(size=(P, q)) * 10
    for name in ("phi", "nu", "pi", "alpha3", "delta", "A", "gamma", "tau"):
        args = (WtW, BtYW, 1.0) if name in ("phi", "nu") else ((np.log(np.ones(K) / K) * n,) if name in ("pi", "alpha3") else ())
        smp.host_update(name, *args)
        t0 = time.perf_counter()
        for _ in range(10): smp.host_update(name, *args)
        print(f"P={P} {name}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms", flush=True)

run(4, 3, 20)
