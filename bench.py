#!/usr/bin/env python
"""bench.py -- Gibbs iterations/sec of the B200 sampler engine on the BASELINE.json workloads.

  python bench.py --gpus N --steps K --warmup W                      (our arm; torchrun launches N ranks)
  python bench.py --workload {nstar,c2,c3,c4,c5} [--scaling weak|strong] ...
  python bench.py --impl reference ...                               (the reference's CPU update functions)

A "step" is ONE full sweep of the reference's warm-start loop (BFMMM_MTT_warm_start order,
BFMMM.h:1500-1554): Z -> pi -> alpha_3 -> Phi -> delta -> A -> gamma -> nu -> tau -> sigma^2 -> chi
-> log-likelihood (covariate-adjusted loop BFMMM.h:3944-4010: + eta, tau_eta, xi and its priors) on
synthetic data of the named shape.  Workloads (BASELINE.json `metric` / `configs`):

  nstar  the metric's shape: functional K=3 P=20 M=3, common 200-point grid, n_funct = 1M   (default)
  c2     configs[1]: the same model with n_funct = 100k
  c3     configs[2]: multivariate K=3 R=64 M=4, n_obs = 1M sharded over the ranks
  c4     configs[3]: covariate-adjusted (eta + xi) D=2 K=3 P=20 M=3, n_funct = 1M, ragged grids n_i ~ U{150..250}
  c5     configs[4]: high-dimensional functional K=4 P=400 (20 x 20 tensor basis, 32 x 32 grid), n_funct = 200k,
         one independent chain per GPU (replicas, no collective)

Functions are sharded across GPUs.  --scaling weak: every rank holds n functions (value = N * steps / time:
sweeps over an n-function shard per second, whole job); --scaling strong: n functions in total.  Defaults:
nstar / c2 weak, c3 / c4 strong (BASELINE states their totals), c5 replicas.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gibbs_iters_per_sec"

WORKLOADS = {
    # kind, K, P, M, D, T, n, default scaling
    "nstar": dict(kind="common", K=3, P=20, M=3, D=0, T=200, n=1_000_000, scaling="weak"),
    "c2": dict(kind="common", K=3, P=20, M=3, D=0, T=200, n=100_000, scaling="weak"),
    "c3": dict(kind="mv", K=3, P=64, M=4, D=0, T=64, n=1_000_000, scaling="strong"),
    "c4": dict(kind="ragged", K=3, P=20, M=3, D=2, T=200, n=1_000_000, scaling="strong"),
    "c5": dict(kind="hd", K=4, P=400, M=3, D=0, T=1024, n=200_000, scaling="replicas"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nstar", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--n", type=int, default=None, help="functions (per GPU when weak, in total when strong)")
    ap.add_argument("--thin", type=int, default=10, help="thinning of the e2e_thinned leg (the reference's thinning_num)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="functions in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Theta_est, ESS and thinned legs")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    if a.n:
        w["n"] = a.n
    if a.scaling and w["scaling"] != "replicas":
        w["scaling"] = a.scaling
    a.w = w
    return a


def shard_sizes(a, world):
    """functions held by each rank"""
    w = a.w
    if w["scaling"] in ("weak", "replicas"):
        return [w["n"]] * world
    base, rem = divmod(w["n"], world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def unit_name(a, world):
    w = a.w
    desc = {"common": f"functional K={w['K']} P={w['P']} M={w['M']}, common {w['T']}-point grid",
            "mv": f"multivariate K={w['K']} R={w['P']} M={w['M']}",
            "ragged": f"covariate-adjusted D={w['D']} K={w['K']} P={w['P']} M={w['M']}, ragged grids (mean n_i = 200)",
            "hd": f"high-dimensional K={w['K']} P={w['P']} (20x20 tensor basis) M={w['M']}, {w['T']}-point grid"}[w["kind"]]
    per = {"weak": f"n_funct={w['n']} per GPU", "strong": f"n_funct={w['n']} in total", "replicas": f"n_funct={w['n']} per chain, one chain per GPU"}
    return f"Gibbs iterations/s (full warm-start sweep, {desc}, {per[w['scaling']]})"


def config_dict(a, world):
    w = a.w
    sizes = shard_sizes(a, world)
    return {"workload": f"{a.workload}: " + unit_name(a, world).split("(", 1)[1].rstrip(")"), "name": a.workload,
            "n_per_gpu": sizes[0], "n_total": sum(sizes) if w["scaling"] != "replicas" else w["n"],
            "K": w["K"], "P": w["P"], "M": w["M"], "D": w["D"], "T": w["T"], "scaling_mode": w["scaling"]}


# ------------------------------------------------------------------------------------------ data
def make_data(a, rank, n, offset):
    """SURVEY.md 8(d) generators; per-rank seed so shards differ, the truth is global."""
    from tests import synth
    w = a.w
    K, P, M, D, T = w["K"], w["P"], w["M"], w["D"], w["T"]
    rng = np.random.default_rng(1 + rank + 1000 * list(WORKLOADS).index(a.workload))
    prng = np.random.default_rng(12345)           # the truth is global: identical on every rank
    par = synth.make_params(prng, K, P, M, D, 0.01)
    pi = prng.dirichlet(np.ones(K))
    Z = np.asfortranarray(rng.dirichlet(10.0 * pi, size=n))
    chi = np.asfortranarray(rng.normal(0, 1, (n, M)))
    X = np.asfortranarray(rng.normal(0, 1, (n, D))) if D else None
    out = dict(par=par, pi=pi, Z=Z, chi=chi, X=X)
    if w["kind"] == "common":
        t = np.linspace(0.0, 1000.0, T)
        ik = synth.equispaced_internal(P, 3)
        B = synth.bspline_design(t, ik, 3)
        y = synth.theta(par, Z, chi) @ B.T
        y += rng.standard_normal((n, T)) * 0.1
        out.update(t=t, ik=ik, B=B, y=y)
    elif w["kind"] == "mv":
        y = synth.theta(par, Z, chi)
        y += rng.standard_normal((n, P)) * 0.1
        out.update(y=np.asfortranarray(y))
    elif w["kind"] == "hd":
        side, ps = int(round(T ** 0.5)), int(round(P ** 0.5))
        g = np.linspace(0.0, 1000.0, side)
        t2 = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
        ik = synth.equispaced_internal(ps, 3)
        B = synth.tensor_design(t2, (ik, ik), 3)
        y = np.empty((n, T))
        for i0 in range(0, n, 20000):            # chunked: theta is n x 400
            sl = slice(i0, min(n, i0 + 20000))
            y[sl] = synth.theta(par, Z[sl], chi[sl]) @ B.T + rng.standard_normal((sl.stop - sl.start, T)) * 0.1
        out.update(t=t2, ik=ik, B=B, y=y)
    else:                                          # ragged grids
        from scipy.interpolate import BSpline
        ik = synth.equispaced_internal(P, 3)
        ni = rng.integers(150, 251, n)
        off = np.concatenate([[0], np.cumsum(ni)]).astype(np.int64)
        N = int(off[-1])
        t = np.empty(N); y = np.empty(N)
        kn = synth.clamped_knots(ik, 3, (0.0, 1000.0))
        for i0 in range(0, n, 50000):              # chunked: the sparse design matrix of 1e7 points is ~0.5 GB
            i1 = min(n, i0 + 50000)
            m = int(off[i1] - off[i0])
            tc = rng.uniform(0, 1000.0, m)
            fn = np.repeat(np.arange(i1 - i0), ni[i0:i1])
            tc = tc[np.lexsort((tc, fn))]           # sorted within each function
            th = synth.theta(par, Z[i0:i1], chi[i0:i1], X[i0:i1])
            Bs = BSpline.design_matrix(tc, kn, 3, extrapolate=False).tocsr()
            yc = np.asarray(Bs.multiply(th[fn]).sum(axis=1)).ravel()
            t[off[i0]:off[i1]] = tc
            y[off[i0]:off[i1]] = yc + rng.standard_normal(m) * 0.1
        out.update(t=t, ik=ik, y=y, off=off)
    return out


def build_engine(a, dat, n, local, offset):
    import bayesfmmm_b200 as bf
    from bayesfmmm_b200 import basis as bfbasis
    from bayesfmmm_b200.engine import FUNCTIONAL, MULTIVARIATE
    w = a.w
    K, P, M, D, T = w["K"], w["P"], w["M"], w["D"], w["T"]
    if w["kind"] == "common":
        eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=dat["y"], T=T, t=dat["t"], degree=3,
                        internal_knots=dat["ik"], boundary=(0.0, 1000.0), device=local, global_offset=offset)
        Pm = bfbasis.pmat_rw1(P)
    elif w["kind"] == "mv":
        eng = bf.Engine(model=MULTIVARIATE, n=n, K=K, P=P, M=M, y=dat["y"], device=local, global_offset=offset)
        Pm = None
    elif w["kind"] == "hd":
        eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=dat["y"], B=dat["B"], T=T, device=local, global_offset=offset)
        Pm = bfbasis.get_P([3, 3], [dat["ik"], dat["ik"]])
    else:
        eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=dat["y"], off=dat["off"], t=dat["t"], degree=3,
                        internal_knots=dat["ik"], boundary=(0.0, 1000.0), X=dat["X"], common_grid=False, device=local,
                        global_offset=offset)
        Pm = bfbasis.pmat_rw1(P)
    return eng, Pm


def algorithmic_bytes(a, n, eng):
    """bytes per launch each pass must move (FP64): rows of the cache + state in / out"""
    w = a.w
    K, M, D = w["K"], w["M"], w["D"]
    if w["kind"] == "ragged":
        bw = eng.dims()[7]
        rows = w["P"] * (1 + bw)                      # least-squares coefficients + band of G_i
        return {"z_kernel": n * (rows + 2 * K + M + D) * 8, "chi_kernel": n * (rows + 1 + K + 2 * M + D) * 8,
                "ssr_kernel": n * (rows + 1 + K + M + D) * 8,
                "stats_kernels": n * (w["P"] + w["P"] * bw + 2 * (K + M + D)) * 8}
    Pc = w["P"]
    return {"z_kernel": n * (Pc + 2 * K + M + D) * 8, "chi_kernel": n * (Pc + 1 + K + 2 * M + D) * 8,
            "ssr_kernel": n * (Pc + 1 + K + M + D) * 8, "stats_kernels": n * (Pc + K + M + D) * 8}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled through NVML every millisecond on a thread."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device):
        self.device, self.sm, self.reasons, self.run, self.err = device, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ent = vis.split(",")[device]
                idx = int(ent) if ent.isdigit() else device
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:
            self.nv, self.err = None, str(exc)

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in self.BAD.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception as exc:
                self.err = str(exc)
                return
            time.sleep(0.001)

    def start(self):
        if self.nv:
            self.run = True
            self.thr = threading.Thread(target=self._loop, daemon=True)
            self.thr.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.err}"], "samples": 0}
        self.run = False
        self.thr.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def bind_to_gpu_numa_node(local):
    """Pins this rank (and so the first touch of its pinned chain slots) to the CPUs of the GPU's NUMA node."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return f"single NUMA node ({len(os.sched_getaffinity(0))} cpus): nothing to bind"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        return f"bound to node {node} ({len(cpus)} cpus) of {len(nodes)}"
    except Exception as exc:
        return f"not bound ({exc})"


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sweep(a, n_sample, repeats=1):
    """Full sweeps of the reference's own update functions (oracle/_ref: Update*.h compiled against the
    shim) -- or of the oracle port when that library is not available -- on n_sample functions of the
    workload's shape.  Returns seconds per sweep and the kind used."""
    from oracle import oracle as orc
    from oracle import ref
    from tests import synth
    w = a.w
    K, P, M, D = w["K"], w["P"], w["M"], w["D"]
    if w["kind"] == "common":
        s = synth.functional_common(seed=3, n=n_sample, T=w["T"], K=K, P=P, M=M)
    elif w["kind"] == "mv":
        s = synth.multivariate(seed=3, n=n_sample, R=P, K=K, M=M)
    elif w["kind"] == "hd":
        s = synth.hd_common(seed=3, n=n_sample, K=K, M=M, side=int(round(w["T"] ** 0.5)), p_side=int(round(P ** 0.5)))
    else:
        s = synth.functional_ragged(seed=3, n=n_sample, K=K, P=P, M=M, D=D, lo=150, hi=250)
    n = s["n"]
    if w["kind"] in ("common", "hd"):
        off = np.arange(n + 1, dtype=np.int64) * s["T"]
        d = orc.Data(n=n, K=K, P=P, M=M, y=s["y"].ravel(), B=np.tile(s["B"], (n, 1)), off=off)
    elif w["kind"] == "mv":
        d = orc.Data(n=n, K=K, P=P, M=M, y=s["y"], identity_basis=True)
    else:
        d = orc.Data(n=n, K=K, P=P, M=M, y=s["y"], B=s["B"], off=s["off"], X=s["X"])
    par = s["par"]
    st = orc.State(nu=par["nu"], Phi=par["Phi"], Z=s["Z"], chi=s["chi"], sigma_sq=0.01, eta=par["eta"], xi=par["xi"])
    rng = np.random.default_rng(0)
    gam = rng.gamma(10000.0 * s["Z"]); u = rng.uniform(size=n); eps = rng.normal(size=(n, M))
    zphi = rng.normal(size=(P, K * M)); znu = rng.normal(size=(P, K))
    zeta = rng.normal(size=(P, max(D, 1) * K)); zxi = rng.normal(size=(P, K * M * max(D, 1)))
    gma = np.ones((K, P, M)); tt = np.ones((K, M)); tau = np.ones(K)
    tau_eta = np.ones((K, max(D, 1))); gxi = np.ones((K, P, max(D, 1), M)); ttxi = np.ones((K, M, max(D, 1)))
    Pm = None if w["kind"] == "mv" else (orc.pmat_rw1(P) if w["kind"] != "hd" else orc.getP([3, 3], [s["internal_knots"]] * 2))
    use_ref = ref.available()
    f = ref if use_ref else orc
    t0 = time.perf_counter()
    for _ in range(repeats):
        f.update_z(d, st, s["pi"], 1.0, 10000.0, gam, u)
        f.update_phi(d, st, gma, tt, zphi)
        f.update_nu(d, st, tau, Pm, znu)
        f.update_sigma(d, st, 1.0, 1.0, 50.0)
        f.update_chi(d, st, eps)
        if D:
            f.update_eta(d, st, tau_eta, Pm, zeta)
            f.update_xi(d, st, gxi, ttxi, zxi)
        f.loglik(d, st)
    dt = (time.perf_counter() - t0) / repeats
    return dt, ("reference" if use_ref else "port")


def default_cpu_sample(a):
    # sized so that one reference sweep takes ~0.5 s on one host core (measured: 2.6 ms per function at the metric's shape)
    return a.cpu_sample or {"nstar": 200, "c2": 200, "c3": 2000, "c4": 100, "c5": 2}[a.workload]


def cpu_baseline(a, unit, n_per_chain):
    n_s = default_cpu_sample(a)
    reps = 5 if a.workload != "c5" else 1
    dt, kind = cpu_reference_sweep(a, n_s, repeats=reps)
    per_iter = dt * (n_per_chain / n_s)            # every reference loop is O(n)
    return {"value": 1.0 / per_iter, "unit": unit, "cores": 1, "kind": kind,
            "sample": (f"{reps} full sweeps (Z, Phi, nu, sigma^2, chi" + (", eta, xi" if a.w["D"] else "") + ", loglik) of the reference's "
                       f"Update*.h functions on {n_s} functions of the workload's shape took {dt:.2f} s each, single-threaded "
                       f"(the reference has no threading); extrapolated linearly in n to {n_per_chain} functions (every "
                       "reference loop is O(n)); " + ("compiled against oracle/shim (Armadillo/Rmath stand-in)" if kind == "reference" else "oracle restatement"))}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    unit = unit_name(a, world)
    sizes = shard_sizes(a, world)
    n_s = default_cpu_sample(a)
    for _ in range(a.warmup):
        cpu_reference_sweep(a, n_s)
    dt, kind = cpu_reference_sweep(a, n_s, repeats=a.steps)
    # whole job on the one thread the reference has: weak / replicas -> N shards (chains) of n functions one after the
    # other, value counts shard-sweeps; strong -> n functions in total
    n_unit = sum(sizes) if a.w["scaling"] == "strong" else sizes[0]
    val = 1.0 / (dt * (n_unit / n_s))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "strong" if a.w["scaling"] == "strong" else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(a, world),
           "cpu_baseline": {"value": val, "unit": unit, "cores": 1, "kind": kind,
                            "sample": (f"each step = one full sweep of the reference's Update*.h functions on a bounded sample of {n_s} "
                                       f"functions of the workload's shape ({dt:.3f} s per step), extrapolated linearly in n to the "
                                       f"job's functions (every reference loop is O(n)); single-threaded: the reference sampler has "
                                       f"no threading (no OpenMP pragmas, P x P BLAS calls only)")},
           "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import bayesfmmm_b200 as bf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sampler engine has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = a.w
    replicas = w["scaling"] == "replicas"
    sizes = shard_sizes(a, world)
    n = sizes[rank]
    offset = 0 if replicas else sum(sizes[:rank])
    n_total = n if replicas else sum(sizes)
    K, P, M, D, T = w["K"], w["P"], w["M"], w["D"], w["T"]
    unit = unit_name(a, world)
    t_gen = time.perf_counter()
    dat = make_data(a, rank, n, offset)
    gen_s = time.perf_counter() - t_gen
    t_create = time.perf_counter()
    eng, Pm = build_engine(a, dat, n, local, offset)
    create_s = time.perf_counter() - t_create
    dat.pop("y", None); dat.pop("t", None)
    eng.set_state(dat["Z"], dat["chi"])
    hyper = bf.default_hyper(True)
    smp = bf.Sampler(eng, hyper=hyper, n_total=n_total, Pmat=Pm, seed=2024 + (rank if replicas else 0))
    par = dat["par"]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=0.01, pi=dat["pi"], alpha3=1.0)
    if D:
        smp.set_cov(eta=par["eta"], xi=par["xi"])
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    if world > 1 and not replicas:
        if w["kind"] == "ragged":                    # totals of sum_i floor(n_i / 2) and sum_i n_i over the shards
            cnt = torch.tensor(list(eng.counts()), dtype=torch.float64, device=dev)
            dist.all_reduce(cnt)
            smp.set_counts(float(cnt[0].item()), float(cnt[1].item()))

        def exchange_id(idb):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(idb), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())

        def allgather(h):
            mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(dev)
            out = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(out, mine)
            return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)
        # One-shot exchange over NVLink peer memory (csrc/p2p_hook.cu), which the engine runs inside the statistics pass's
        # final reduction and inside the SSR pass (no kernels of its own; 2 GPUs: 0.287 ms per sweep, 4 GPUs: 0.295).
        # BFMMM_NCCL_ALLREDUCE=1 selects the native NCCL hook instead (csrc/nccl_hook.cu); so does a failed peer mapping
        # on any rank.
        use_p2p = not os.environ.get("BFMMM_NCCL_ALLREDUCE")
        ok = torch.ones(1, device=dev)
        if use_p2p:
            try:
                smp.enable_p2p(rank, world, eng.stats_buffer()[1], allgather)
            except Exception as exc:          # no peer access between some pair of devices
                print(f"rank {rank}: peer-memory all-reduce unavailable ({exc}); using NCCL", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not use_p2p or ok.item() == 0:
            smp.enable_nccl(rank, world, exchange_id)
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn(steps)
        e1.record(ext)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    units_per_step = world if w["scaling"] in ("weak", "replicas") else 1     # shard-sweeps (weak) / whole-data sweeps (strong)

    # ---- headline: K sweeps with everything resident in HBM
    smp.run(bf.SWEEP_FULL, a.warmup)
    prof0 = smp.profile()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = eng.launch_count
    ms = timed(lambda k: smp.run(bf.SWEEP_FULL, k), a.steps)
    launches = eng.launch_count - l0
    prof1 = smp.profile()
    value = units_per_step * a.steps / (ms * 1e-3)

    # ---- e2e: the same sweep through the host-buffer API, the new Z and chi copied into the caller's chain storage
    # (the reference's chain containers hold every iteration: BFMMM.h:1211-1218, slice i+1 written by every update).
    # Page-locked chain slots, two, used alternately: slice i travels while sweep i+1 runs.
    slots = [(torch.empty((K, n), dtype=torch.float64).pin_memory().numpy().T,      # column-major n x K view
              torch.empty((M, n), dtype=torch.float64).pin_memory().numpy().T) for _ in range(2)]
    _, stats_len = eng.stats_buffer()

    def e2e_steps(k, thin=1):
        for it in range(k):
            smp.step(bf.SWEEP_FULL)
            if (it + 1) % thin == 0:
                eng.get_state_begin(*slots[(it // thin) % 2])
        eng.get_state_wait()
    e2e_steps(2)
    e_steps = max(20, a.steps)
    ms_e = timed(e2e_steps, e_steps)
    clk = clocks.stop()
    e2e_value = units_per_step * e_steps / (ms_e * 1e-3)
    n_push = 3 if not D else 4
    h2d = n_push * P * (K * (M + 1) * (1 + D)) * 8       # pushes of the global coefficients per sweep
    d2h = n * (K + M) * 8 + 3 * stats_len * 8            # Z and chi into the chain + the statistics read-backs
    e2e = {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e_steps, "ms_per_step": ms_e / e_steps,
           "delivery": "every sweep's Z and chi into pinned host chain slots (what *_Theta_est's chain containers hold)"}
    out = {"metric": METRIC, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms / a.steps, "higher_is_better": True,
           "scaling": "strong" if w["scaling"] == "strong" else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "gpu_launches": int(launches)}
    cfg = config_dict(a, world)
    cfg.update({"timing": "inputs larger than L2 (cache + state per pass vs 126 MB L2)" if n * (P + K + M) * 8 > 126e6 else
                          "L2 flushed by the sweep itself: four passes over cache + state, inputs partly L2 resident at this n",
                "rng": "device Philox (no injected draws)", "create_s_untimed": create_s, "datagen_s_untimed": gen_s,
                "numa": numa})
    out["config"] = cfg
    out["clocks"] = clk
    out["host_split_ms_per_step"] = {k.replace("_s", ""): (prof1[k] - prof0[k]) / a.steps * 1e3 for k in prof1}
    out["e2e"] = e2e

    if not a.no_extras:
        # ---- thinned delivery (the warm-start drivers store every thinning_num-th draw: BFMMM.h:1695-1718)
        e2e_steps(a.thin, a.thin)
        t_steps = max(a.thin * 4, (e_steps // a.thin) * a.thin)
        ms_t = timed(lambda k: e2e_steps(k, a.thin), t_steps)
        out["e2e_thinned"] = {"value": units_per_step * t_steps / (ms_t * 1e-3), "unit": unit, "thinning_num": a.thin,
                              "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n * (K + M) * 8 / a.thin + 3 * stats_len * 8),
                              "steps": t_steps, "ms_per_step": ms_t / t_steps,
                              "delivery": f"Z and chi of every {a.thin}-th sweep (the stored draws of the warm-start drivers)"}
        # ---- the Theta_est sweep (BFMMM_Theta, BFMMM.h:1253-1298: Phi, delta, A, gamma, tau, sigma^2, chi, loglik; Z and nu fixed)
        smp.run(bf.SWEEP_THETA, 3)
        ms_theta = timed(lambda k: smp.run(bf.SWEEP_THETA, k), a.steps)
        out["theta_est_sweep"] = {"value": units_per_step * a.steps / (ms_theta * 1e-3), "unit": "Theta_est sweeps/s (Z, nu fixed)",
                                  "ms_per_step": ms_theta / a.steps}
        # ---- ESS/sec of Z (BASELINE.json names it; the reference has no ESS code): batch-means effective
        # sample size of each Z_ik chain for 256 monitored functions over a further `ess_steps` sweeps,
        # median over (i, k), per second of sweep time
        ess_steps, nmon = (400 if a.workload in ("nstar", "c2", "c3") else 100), min(256, n)
        zc = np.zeros((ess_steps, nmon, K))
        barrier()
        t_e0 = time.perf_counter()
        for it in range(ess_steps):
            smp.step(bf.SWEEP_FULL)
            zc[it] = eng.get_state_rows(0, nmon, chi=False)[0]
        torch.cuda.synchronize()
        t_ess = time.perf_counter() - t_e0
        nb = 20
        bm = zc.reshape(nb, ess_steps // nb, nmon, K).mean(axis=1)
        var_chain = zc.var(axis=0, ddof=1)
        var_bm = bm.var(axis=0, ddof=1) * (ess_steps // nb)
        ess = np.where(var_bm > 0, ess_steps * var_chain / np.maximum(var_bm, 1e-300), float(ess_steps))
        out["ess_z"] = {"median_ess": float(np.median(ess)), "sweeps": ess_steps, "monitored_functions": nmon,
                        "ess_per_sec": float(np.median(ess) / t_ess), "method": "batch means (20 batches), median over (i,k)",
                        "mean_accept_rate": smp.last_accept / n_total}

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline
    def kernel_ms(fn, reps=20):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    g = smp.get()
    pi_now, a3 = g["pi"], g["alpha3"]
    ab = algorithmic_bytes(a, n, eng)
    t_z = kernel_ms(lambda: eng.update_z_async(pi_now, a3, hyper.a_Z_PM))
    if eng.debug_z_propose(pi_now, a3, hyper.a_Z_PM):
        # the Z step is two kernels: the proposal (reads Z, writes z*, lr, lu; in the sweep it runs ahead of time on a
        # side stream while the host draws the Gaussian blocks) and the pass over the cache that accepts
        t_p = kernel_ms(lambda: eng.debug_z_propose(pi_now, a3, hyper.a_Z_PM))
        kern = {"z_propose_kernel": t_p, "z_kernel": t_z - t_p}
        ab["z_propose_kernel"] = n * 8 * (2 * K + 2)
        ab["z_kernel"] += n * 8 * (K + 2)
    else:
        kern = {"z_kernel": t_z}
    t_ssr = kernel_ms(lambda: eng.ssr_async())
    if eng.debug_moments_valid():
        # common basis, no covariates: updateSigma's pass leaves the per-function moments and updateChi draws from them
        # (csrc/moments_kernels.cu) -- the sweep's two kernels are these, not ssr_kernel + chi_kernel
        n8 = n * 8
        kern["moments_kernel"] = t_ssr
        kern["chi_draw_kernel"] = kernel_ms(lambda: eng.update_chi_async())
        ab["moments_kernel"] = n8 * (w["P"] + 1 + K + M + (M + 1))
        ab["chi_draw_kernel"] = n8 * (K + M + (M + 1) + M)
        del ab["ssr_kernel"], ab["chi_kernel"]
    else:
        kern["ssr_kernel"] = t_ssr
        kern["chi_kernel"] = kernel_ms(lambda: eng.update_chi_async())
    kern["stats_kernels"] = kernel_ms(lambda: eng.suffstats_async())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kinfo = {k: {"ms": v, "algorithmic_bytes": ab[k], "gbs": ab[k] / (v * 1e-3) / 1e9, "frac": ab[k] / (v * 1e-3) / 1e9 / peak}
             for k, v in kern.items()}
    dom = max(kinfo, key=lambda k: kinfo[k]["ms"])
    # dram bytes of the same kernel from the committed `ncu --set full` capture (profiles/r02_traffic.json), used only
    # when that capture was taken on a build of exactly these device-code sources (md5 over csrc/ minus the host-only
    # translation units, bayesfmmm_b200/_lib.py:kernel_source_hash) and this shard size
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tr.get(a.workload, {})
        if ent.get("n_per_gpu") == n and ent.get("source_md5") == bf._lib.kernel_source_hash():
            traffic = ent["kernels"].get(dom)
            traffic_src = "profiles/r02_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, same device-code sources: md5 over csrc/ minus the host-only translation units)"
        elif ent:
            traffic_src = "profiles/r02_traffic.json was captured on another build or shard size: not reported"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kinfo[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kinfo[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kinfo[dom]["algorithmic_bytes"], "kernels": kinfo,
                # covariate-adjusted sweeps take the statistics and the SSR a second time after chi (BFMMM.h:3976-4000)
                "device_ms_per_step_sum_of_kernels": sum(v["ms"] for v in kinfo.values()) + ((kinfo["stats_kernels"]["ms"] + kinfo["ssr_kernel"]["ms"]) if D else 0.0),
                "sweep_algorithmic_gbs": sum(ab.values()) / (ms / a.steps * 1e-3) / 1e9}
    if w["kind"] == "ragged":
        # the pair cross-Gram of the ragged statistics is FP64 tensor-pipe work: npairs x bw P multiply-adds per function
        bw = eng.dims()[7]
        q = K * (1 + D) * (1 + M)
        flops = 2.0 * n * (q * (q + 1) // 2) * bw * P
        dmma_peak = 148 * 4 * 1.965e9 / 16.1 * 512 / 1e12    # tools/micro/dmma_dfma.cu: 16.1 cycles per mma.m8n8k4.f64 per sub-partition
        roofline["ragged_cross_gram"] = {"bound": "tensor", "flops_per_launch": flops, "unit": "TFLOP/s", "peak": dmma_peak,
                                         "peak_source": "tools/micro/dmma_dfma.cu (FP64 DMMA issue rate measured on B200)",
                                         "achieved_lower_bound": flops / (kinfo["stats_kernels"]["ms"] * 1e-3) / 1e12,
                                         "note": "stats_kernels' time also contains the W'W / B'Y'W pass"}
    out["roofline"] = roofline
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(a, unit, n)
            except Exception as exc:    # keep the GPU line even if the CPU checker cannot run
                out["cpu_baseline"] = {"value": None, "unit": unit, "cores": 1, "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(out))
    barrier()                 # no rank unmaps its peers' mailboxes while another is still exchanging
    smp.close(); eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
