#!/usr/bin/env python
"""bench.py -- Gibbs iterations/sec of the B200 sampler engine on the BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun launches N ranks)
  python bench.py --impl reference ...                     (the reference's CPU update functions)

A "step" is ONE full sweep of the reference's warm-start loop (BFMMM_MTT_warm_start order,
BFMMM.h:1500-1554): Z -> pi -> alpha_3 -> Phi -> delta -> A -> gamma -> nu -> tau -> sigma^2 -> chi
-> log-likelihood, on synthetic data of the named shape (n_funct=1M, K=3, P=20, M=3, common 200-point
grid).  Functions are sharded across GPUs; each rank holds n_funct functions ("scaling": "weak"),
so `value` = N * steps / time = sweeps over a 1M-function shard per second, whole job.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gibbs_iters_per_sec"
UNIT = "Gibbs iterations/s (full warm-start sweep, n_funct=1M K=3 P=20 M=3 per GPU)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="functions per GPU")
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--P", type=int, default=20)
    ap.add_argument("--M", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=2000, help="functions in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"functional FMMM K={a.K} P={a.P} M={a.M} n_funct={a.n} per GPU, common {a.T}-point grid, "
            f"full warm-start sweep (BFMMM_MTT_warm_start order)")


# ------------------------------------------------------------------------------------------ data
def make_data(a, rank):
    """SURVEY.md 8(d) generator; per-rank seed so shards differ."""
    from tests import synth
    rng = np.random.default_rng(1 + rank)
    T, P, K, M, n = a.T, a.P, a.K, a.M, a.n
    t = np.linspace(0.0, 1000.0, T)
    ik = synth.equispaced_internal(P, 3)
    B = synth.bspline_design(t, ik, 3)
    prng = np.random.default_rng(12345)           # the truth is global: identical on every rank
    par = synth.make_params(prng, K, P, M, 0, 0.01)
    pi = prng.dirichlet(np.ones(K))
    Z = np.asfortranarray(rng.dirichlet(10.0 * pi, size=n))
    chi = np.asfortranarray(rng.normal(0, 1, (n, M)))
    th = synth.theta(par, Z, chi)
    y = th @ B.T
    y += rng.standard_normal((n, T)) * 0.1
    return dict(t=t, ik=ik, B=B, par=par, pi=pi, Z=Z, chi=chi, y=y)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sweep(a, n_sample, repeats=1):
    """One full sweep of the reference's own update functions (oracle/_ref: Update*.h compiled against the
    shim) -- or of the oracle port when that library is not available -- on n_sample functions of the
    benchmark shape.  Returns seconds per sweep and the kind used."""
    from oracle import oracle as orc
    from oracle import ref
    from tests import synth
    s = synth.functional_common(seed=3, n=n_sample, T=a.T, K=a.K, P=a.P, M=a.M)
    n, T = s["n"], s["T"]
    off = np.arange(n + 1, dtype=np.int64) * T
    d = orc.Data(n=n, K=a.K, P=a.P, M=a.M, y=s["y"].ravel(), B=np.tile(s["B"], (n, 1)), off=off)
    par = s["par"]
    st = orc.State(nu=par["nu"], Phi=par["Phi"], Z=s["Z"], chi=s["chi"], sigma_sq=0.01)
    rng = np.random.default_rng(0)
    K, P, M = a.K, a.P, a.M
    gam = rng.gamma(10000.0 * s["Z"]); u = rng.uniform(size=n); eps = rng.normal(size=(n, M))
    zphi = rng.normal(size=(P, K * M)); znu = rng.normal(size=(P, K))
    gma = np.ones((K, P, M)); tt = np.ones((K, M)); tau = np.ones(K)
    Pm = orc.pmat_rw1(P)
    use_ref = ref.available()
    impl = ref if use_ref else orc
    t0 = time.perf_counter()
    for _ in range(repeats):
        if use_ref:
            ref.update_z(d, st, s["pi"], 1.0, 10000.0, gam, u)
            ref.update_phi(d, st, gma, tt, zphi)
            ref.update_nu(d, st, tau, Pm, znu)
            ref.update_sigma(d, st, 1.0, 1.0, 50.0)
            ref.update_chi(d, st, eps)
            ref.loglik(d, st)
        else:
            orc.update_z(d, st, s["pi"], 1.0, 10000.0, gam, u)
            orc.update_phi(d, st, gma, tt, zphi)
            orc.update_nu(d, st, tau, Pm, znu)
            orc.update_sigma(d, st, 1.0, 1.0, 50.0)
            orc.update_chi(d, st, eps)
            orc.loglik(d, st)
    dt = (time.perf_counter() - t0) / repeats
    return dt, ("reference" if use_ref else "port")


def cpu_baseline(a):
    n_s = a.cpu_sample
    dt, kind = cpu_reference_sweep(a, n_s)
    per_iter = dt * (a.n / n_s)            # every reference loop is O(n)
    return {"value": 1.0 / per_iter, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": (f"one full sweep (Z, Phi, nu, sigma^2, chi, loglik) of the reference's Update*.h functions "
                       f"on {n_s} functions of the benchmark shape took {dt:.2f} s single-threaded; extrapolated "
                       f"linearly in n to {a.n} functions (every reference loop is O(n)); "
                       + ("compiled against oracle/shim (Armadillo/Rmath stand-in)" if kind == "reference" else "oracle restatement"))}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 3))
    for _ in range(min(a.warmup, 1)):
        cpu_reference_sweep(a, max(10, a.cpu_sample // 10))
    dt, kind = cpu_reference_sweep(a, a.cpu_sample, repeats=steps)
    per_iter = dt * (a.n / a.cpu_sample)
    val = 1.0 / per_iter
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
           "warmup": min(a.warmup, 1), "ms_per_step": per_iter * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(a)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                            "sample": (f"{steps} sweeps on {a.cpu_sample} functions of the benchmark shape, {dt:.2f} s each, "
                                       f"extrapolated linearly to n={a.n}; the reference sampler is single-threaded "
                                       f"(no OpenMP pragmas, P x P BLAS calls only)")},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import bayesfmmm_b200 as bf
    from bayesfmmm_b200.engine import FUNCTIONAL
    from bayesfmmm_b200 import basis as bfbasis

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sampler engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dat = make_data(a, rank)
    n, K, P, M, T = a.n, a.K, a.P, a.M, a.T
    t_create = time.perf_counter()
    eng = bf.Engine(model=FUNCTIONAL, n=n, K=K, P=P, M=M, y=dat["y"], T=T, t=dat["t"], degree=3,
                    internal_knots=dat["ik"], boundary=(0.0, 1000.0), device=local, global_offset=rank * n)
    create_s = time.perf_counter() - t_create
    del dat["y"]
    eng.set_state(dat["Z"], dat["chi"])
    hyper = bf.default_hyper(True)
    smp = bf.Sampler(eng, hyper=hyper, n_total=n * world, Pmat=bfbasis.pmat_rw1(P), seed=2024)
    par = dat["par"]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=0.01, pi=dat["pi"], alpha3=1.0)
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    if world > 1:
        # native NCCL all-reduce of the statistics buffer on the engine's stream (csrc/nccl_hook.cu); rank 0's
        # unique id travels through the torch.distributed process group.  BFMMM_PY_ALLREDUCE=1 selects the
        # generic hook through torch.distributed instead (what a caller without NCCL handles would plug in).
        if os.environ.get("BFMMM_PY_ALLREDUCE"):
            class _Buf:   # zero-copy torch view of the engine's statistics buffer
                def __init__(self, ptr, ln):
                    self.__cuda_array_interface__ = {"shape": (ln,), "typestr": "<f8", "data": (ptr, False), "version": 3}
            ptr, ln = eng.stats_buffer()
            stats_t = torch.as_tensor(_Buf(ptr, ln), device=torch.device("cuda", local))
            torch.cuda.set_stream(ext)          # NCCL work is ordered on the engine's stream

            def allreduce(p, l, stream):        # (device pointer, doubles): the whole buffer or one slot of it
                off = (p - ptr) // 8
                dist.all_reduce(stats_t[off:off + l])
            smp.set_allreduce(allreduce)
        else:
            def exchange_id(idb):
                t = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", local))
                if rank == 0:
                    t.copy_(torch.frombuffer(bytearray(idb), dtype=torch.uint8))
                dist.broadcast(t, 0)
                return bytes(t.cpu().numpy().tobytes())

            def allgather(h):
                mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(torch.device("cuda", local))
                out = [torch.zeros(64, dtype=torch.uint8, device=torch.device("cuda", local)) for _ in range(world)]
                dist.all_gather(out, mine)
                return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)
            # From 4 ranks on: one-shot all-reduce over NVLink peer memory (csrc/p2p_hook.cu; measured at 8 ranks
            # 0.399 ms per sweep against 0.415 ms with ncclAllReduce; at 2 ranks 0.357 against 0.348, so NCCL stays
            # there).  BFMMM_NCCL_ALLREDUCE=1 / BFMMM_P2P_ALLREDUCE=1 force one or the other; a failed peer
            # mapping on any rank selects the native NCCL hook (csrc/nccl_hook.cu).
            use_p2p = (world >= 4 or bool(os.environ.get("BFMMM_P2P_ALLREDUCE"))) and not os.environ.get("BFMMM_NCCL_ALLREDUCE")
            ok = torch.ones(1, device=torch.device("cuda", local))
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

    # ---- headline: K sweeps with everything resident in HBM
    smp.run(bf.SWEEP_FULL, a.warmup)
    prof0 = smp.profile()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = eng.launch_count
    ms = timed(lambda k: smp.run(bf.SWEEP_FULL, k), a.steps)
    launches = eng.launch_count - l0
    clk = clocks.stop()
    prof1 = smp.profile()
    value = world * a.steps / (ms * 1e-3)

    # ---- the Theta_est sweep (BFMMM_Theta, BFMMM.h:1253-1298: Phi, delta, A, gamma, tau, sigma^2, chi, loglik; Z and nu fixed)
    smp.run(bf.SWEEP_THETA, 3)
    ms_theta = timed(lambda k: smp.run(bf.SWEEP_THETA, k), a.steps)
    theta_est = {"value": world * a.steps / (ms_theta * 1e-3), "unit": "Theta_est sweeps/s (Z, nu fixed)", "ms_per_step": ms_theta / a.steps}

    # ---- e2e: the same sweep through the host-buffer API, copying the new Z and chi back into the
    # caller's chain storage every iteration (what the reference's chain containers require)
    # page-locked chain slots (two, used alternately: slice i travels while sweep i+1 runs)
    slots = [(torch.empty((K, n), dtype=torch.float64).pin_memory().numpy().T,      # column-major n x K view
              torch.empty((M, n), dtype=torch.float64).pin_memory().numpy().T) for _ in range(2)]
    _, stats_len = eng.stats_buffer()

    def e2e_steps(k):
        for it in range(k):
            smp.step(bf.SWEEP_FULL)
            eng.get_state_begin(*slots[it % 2])
        eng.get_state_wait()
    e2e_steps(2)
    e_steps = max(5, a.steps // 5)
    ms_e = timed(e2e_steps, e_steps)
    e2e_value = world * e_steps / (ms_e * 1e-3)
    h2d = 3 * P * (K * (M + 1)) * 8                      # three pushes of the global coefficients per sweep
    d2h = n * (K + M) * 8 + 3 * stats_len * 8            # Z and chi into the chain + three statistics read-backs

    # ---- ESS/sec of Z (BASELINE.json names it; the reference has no ESS code): batch-means effective
    # sample size of each Z_ik chain for 256 monitored functions over a further `ess_steps` sweeps,
    # median over (i, k), per second of sweep time
    ess_steps, nmon = 400, 256
    zc = np.zeros((ess_steps, nmon, K))
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
    ess_z = {"median_ess": float(np.median(ess)), "sweeps": ess_steps, "monitored_functions": nmon,
             "ess_per_sec": float(np.median(ess) / t_ess), "method": "batch means (20 batches), median over (i,k)",
             "mean_accept_rate": smp.last_accept / (n * world)}

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
    kern = {
        "z_kernel": (kernel_ms(lambda: eng.update_z_async(pi_now, a3, hyper.a_Z_PM)), n * (P + 4 * K + M) * 8),   # Z and log Z read + written
        "chi_kernel": (kernel_ms(lambda: eng.update_chi_async()), n * (P + 1 + K + 2 * M) * 8),
        "ssr_kernel": (kernel_ms(lambda: eng.ssr_async()), n * (P + 1 + K + M) * 8),
        "stats_kernel": (kernel_ms(lambda: eng.suffstats_async()), n * (P + K + M) * 8),
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # raw_stream_floor_ms: the HBM lower bound of the REFERENCE formulation of the same pass (stream the n x T
    # observations) -- information only: frac is scored on the bytes this engine's kernels move
    raw_bytes = {"z_kernel": n * (T + 2 * K + M) * 8, "chi_kernel": n * (T + K + 2 * M) * 8,
                 "ssr_kernel": n * (T + K + M) * 8, "stats_kernel": n * (T + K + M) * 8}
    kinfo = {k: {"ms": v[0], "algorithmic_bytes": v[1], "gbs": v[1] / (v[0] * 1e-3) / 1e9,
                 "frac": v[1] / (v[0] * 1e-3) / 1e9 / peak,
                 "raw_stream_floor_ms": raw_bytes[k] / (peak * 1e9) * 1e3} for k, v in kern.items()}
    dom = max(kinfo, key=lambda k: kinfo[k]["ms"])
    traffic = None
    try:   # dram bytes of the same kernel from the committed `ncu --set full` capture (profiles/), same n
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tr.get("n_per_gpu") == n:
            traffic = tr["kernels"].get(dom, tr["kernels"].get(dom + "_tma"))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kinfo[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kinfo[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kinfo[dom]["algorithmic_bytes"], "kernels": kinfo,
                "device_ms_per_step_sum_of_kernels": sum(v["ms"] for v in kinfo.values())}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(a), "n_per_gpu": n, "n_total": n * world, "K": K, "P": P, "M": M, "T": T,
                      "timing": "inputs larger than L2 (208-280 MB of projected cache + state per pass vs 126 MB L2)",
                      "rng": "device Philox (no injected draws)", "create_s_untimed": create_s},
           "clocks": clk, "gpu_launches": int(launches),
           "host_split_ms_per_step": {k.replace("_s", ""): (prof1[k] - prof0[k]) / a.steps * 1e3 for k in prof1},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "steps": e_steps, "ms_per_step": ms_e / e_steps},
           "theta_est_sweep": theta_est,
           "ess_z": ess_z,
           "roofline": roofline}
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(a)
            except Exception as exc:    # keep the GPU line even if the CPU checker cannot run
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {exc}"}
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
