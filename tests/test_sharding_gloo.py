"""N>1 path on the CPU (world_size 2, gloo): the sharding protocol of DESIGN.md section 4.
Each rank holds a contiguous block of functions, forms its shard's sufficient statistics (NumPy
stands in for the device kernels here), the statistics buffer is summed with an all-reduce, and
every rank then runs the product's host updates (detached Sampler) with the same Philox seed.
Checks: (1) all ranks end with bit-identical globals, (2) they equal the single-shard result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bayesfmmm_b200 as bf
from oracle import oracle as orc
from tests import cases


def _features(Z, chi, K, M):
    cols = []
    for k in range(K):
        for mm in range(M + 1):
            w = Z[:, k].copy()
            if mm:
                w = w * chi[:, mm - 1]
            cols.append(w)
    return np.stack(cols, axis=1)


def _host_sweep(s, d, WtW, BtYW, ssr, slz, n_total, seed=99):
    K, P, M = d.K, d.P, d.M
    G = s["B"].T @ s["B"]
    smp = bf.Sampler(hyper=bf.default_hyper(True), n_total=n_total, Pmat=orc.pmat_rw1(P), seed=seed,
                     dims=(d.n, K, P, M, 0, 0), G=G, sum_half_total=n_total * (s["T"] // 2),
                     n_points_total=n_total * s["T"])
    par = s["par"]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=par["sigma_sq"], pi=s["pi"], alpha3=1.2)
    smp.host_update("pi", slz)
    smp.host_update("alpha3", slz)
    smp.host_update("phi", WtW, BtYW, 1.0)
    smp.host_update("delta"); smp.host_update("A"); smp.host_update("gamma")
    smp.host_update("nu", WtW, BtYW, 1.0)
    smp.host_update("tau")
    smp.host_update("sigma", ssr, 1.0, False)
    g = smp.get()
    smp.close()
    return np.concatenate([g["nu"].ravel(), g["Phi"].ravel(), g["pi"], [g["alpha3"], g["sigma_sq"]], g["tau"],
                           g["delta"].ravel(), g["gamma"].ravel(), g["A"].ravel()])


def _shard_stats(s, d, lo, hi):
    K, M = d.K, d.M
    Z, chi = s["Z"][lo:hi], s["chi"][lo:hi]
    W = _features(Z, chi, K, M)
    BtY = s["y"][lo:hi] @ s["B"]
    th = W @ np.stack([np.concatenate([[s["par"]["nu"][k]], s["par"]["Phi"][k].T]) for k in range(K)]).reshape(-1, d.P)
    ssr = float(((s["y"][lo:hi] - th @ s["B"].T) ** 2).sum())
    return W.T @ W, BtY.T @ W, ssr, np.log(Z).sum(axis=0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, d, st = cases.build("F_common")
    n = d.n
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    WtW, BtYW, ssr, slz = _shard_stats(s, d, lo, hi)
    buf = torch.from_numpy(np.concatenate([WtW.ravel(order="F"), BtYW.ravel(order="F"), [ssr], slz]))
    dist.all_reduce(buf)                               # the only exchange of the multi-GPU path
    b = buf.numpy()
    q2 = d.K * (d.M + 1)
    WtW = b[:q2 * q2].reshape((q2, q2), order="F")
    BtYW = b[q2 * q2:q2 * q2 + d.P * q2].reshape((d.P, q2), order="F")
    ssr = float(b[q2 * q2 + d.P * q2]); slz = b[q2 * q2 + d.P * q2 + 1:]
    q.put((rank, _host_sweep(s, d, WtW, BtYW, ssr, slz, n)))
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank_globals():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0], res[1])              # identical globals on every rank, no broadcast
    s, d, st = cases.build("F_common")
    WtW, BtYW, ssr, slz = _shard_stats(s, d, 0, d.n)
    single = _host_sweep(s, d, WtW, BtYW, ssr, slz, d.n)
    assert np.max(np.abs(res[0] - single) / (1 + np.abs(single))) < 1e-9
