"""Host-side mirror of include/bfmmm_sampler.h: the driver loops (BFMMM_Theta / BFMMM_Nu_Z /
BFMMM_MTT_warm_start order) and the individual host updates."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import dp, load_library
from .engine import Engine, EngineError

SWEEP_THETA, SWEEP_NU_Z, SWEEP_FULL = 0, 1, 2


class Hyper(C.Structure):
    _fields_ = [("c", C.c_double * 8), ("b", C.c_double), ("nu_1", C.c_double),
                ("alpha1l", C.c_double), ("alpha2l", C.c_double), ("beta1l", C.c_double), ("beta2l", C.c_double),
                ("a_Z_PM", C.c_double), ("a_pi_PM", C.c_double), ("var_alpha3", C.c_double),
                ("var_epsilon1", C.c_double), ("var_epsilon2", C.c_double),
                ("alpha_nu", C.c_double), ("beta_nu", C.c_double), ("alpha_eta", C.c_double), ("beta_eta", C.c_double),
                ("alpha_0", C.c_double), ("beta_0", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


def default_hyper(theta_est=True, **over):
    h = Hyper()
    load_library().bfmmm_hyper_defaults(C.byref(h), 1 if theta_est else 0)
    for k, v in over.items():
        if k == "c":
            for i, x in enumerate(v):
                h.c[i] = float(x)
        else:
            setattr(h, k, float(v))
    return h


def _f(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O", "F"])


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


def _cubes(x, K):
    """K cubes (k, P, D, M) -> back-to-back column-major storage."""
    return np.ascontiguousarray(np.stack([np.asfortranarray(x[k]).ravel(order="F") for k in range(K)]))


class Sampler:
    def __init__(self, engine: Engine = None, *, hyper: Hyper = None, n_total: int = 0, Pmat=None, seed: int = 1,
                 dims=None, G=None, sum_half_total: float = 0.0, n_points_total: float = 0.0):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.hyper = hyper if hyper is not None else default_hyper()
        Pm = _f(Pmat) if Pmat is not None else None
        if engine is not None:
            self.engine = engine
            self.K, self.P, self.M, self.D = engine.K, engine.P, engine.M, engine.D
            rc = self._lib.bfmmm_sampler_create(engine._h, C.byref(self.hyper), C.c_int64(n_total), _p(Pm),
                                                C.c_uint64(seed), C.byref(self._h))
        else:
            self.engine = None
            dims = tuple(dims) + (0, 0) if len(dims) == 6 else tuple(dims)
            n, K, P, M, D, model, ragged, bw = dims
            self.K, self.P, self.M, self.D = K, P, M, D
            d = (C.c_int32 * 8)(*dims)
            Gf = _f(G) if G is not None else _f(np.eye(P))
            rc = self._lib.bfmmm_sampler_create_detached(d, C.byref(self.hyper), C.c_int64(n_total), _p(Pm), _p(Gf),
                                                         C.c_double(sum_half_total), C.c_double(n_points_total),
                                                         C.c_uint64(seed), C.byref(self._h))
        self._chk(rc)
        self._cb = None

    def _chk(self, rc):
        if rc != 0:
            raise EngineError(self._lib.bfmmm_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bfmmm_sampler_destroy(self._h)
            self._h = C.c_void_p()
        if getattr(self, "_p2p", None) is not None and self._p2p.value:
            self._lib.bfmmm_p2p_destroy(self._p2p)
            self._p2p = C.c_void_p()
        if getattr(self, "_nccl", None) is not None and self._nccl.value:
            self._lib.bfmmm_nccl_destroy(self._nccl)
            self._nccl = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def set(self, nu=None, Phi=None, sigma_sq=None, pi=None, alpha3=None, delta=None, gamma=None, A=None, tau=None):
        arrs = [None if a is None else _f(a) for a in (nu, Phi)]
        s2 = None if sigma_sq is None else C.byref(C.c_double(sigma_sq))
        a3 = None if alpha3 is None else C.byref(C.c_double(alpha3))
        rest = [None if a is None else _f(a) for a in (pi, delta, gamma, A, tau)]
        self._chk(self._lib.bfmmm_sampler_set(self._h, _p(arrs[0]), _p(arrs[1]), s2, _p(rest[0]), a3, _p(rest[1]),
                                              _p(rest[2]), _p(rest[3]), _p(rest[4])))

    def get(self):
        K, P, M = self.K, self.P, self.M
        nu = np.zeros((K, P), order="F"); Phi = np.zeros((K, P, M), order="F"); pi = np.zeros(K)
        delta = np.zeros((K, M), order="F"); gamma = np.zeros((K, P, M), order="F"); A = np.zeros((K, 2), order="F")
        tau = np.zeros(K)
        s2, a3, ll = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._lib.bfmmm_sampler_get(self._h, _p(nu), _p(Phi), C.byref(s2), _p(pi), C.byref(a3), _p(delta),
                                              _p(gamma), _p(A), _p(tau), C.byref(ll)))
        return dict(nu=nu, Phi=Phi, sigma_sq=s2.value, pi=pi, alpha3=a3.value, delta=delta, gamma=gamma, A=A, tau=tau,
                    loglik=ll.value)

    def set_cov(self, eta=None, xi=None, tau_eta=None, delta_xi=None, gamma_xi=None, A_xi=None):
        K = self.K
        eta_f = None if eta is None else _f(eta)
        xi_f = None if xi is None else _cubes(xi, K)
        te = None if tau_eta is None else _f(tau_eta)
        dx = None if delta_xi is None else _f(delta_xi)
        gx = None if gamma_xi is None else _cubes(gamma_xi, K)
        ax = None if A_xi is None else _f(A_xi)
        self._chk(self._lib.bfmmm_sampler_set_cov(self._h, _p(eta_f), _p(xi_f), _p(te), _p(dx), _p(gx), _p(ax)))

    def get_cov(self):
        K, P, M, D = self.K, self.P, self.M, self.D
        eta = np.zeros((P, D, K), order="F"); xi = np.zeros((K, P * D * M)); te = np.zeros((K, D), order="F")
        dx = np.zeros((K, M, D), order="F"); gx = np.zeros((K, P * D * M)); ax = np.zeros((K, 2, D), order="F")
        self._chk(self._lib.bfmmm_sampler_get_cov(self._h, _p(eta), _p(xi), _p(te), _p(dx), _p(gx), _p(ax)))
        unp = lambda a: np.stack([a[k].reshape((P, D, M), order="F") for k in range(K)])
        return dict(eta=eta, xi=unp(xi), tau_eta=te, delta_xi=dx, gamma_xi=unp(gx), A_xi=ax)

    # ------------------------------------------------------------------ loops
    def step(self, sweep=SWEEP_FULL, beta=1.0):
        self._chk(self._lib.bfmmm_sampler_step(self._h, int(sweep), C.c_double(beta)))

    def run(self, sweep, n_iter):
        self._chk(self._lib.bfmmm_sampler_run(self._h, int(sweep), int(n_iter)))

    def tempered_transition(self, N_t, beta_N_t):
        logA, acc = C.c_double(), C.c_int()
        self._chk(self._lib.bfmmm_sampler_tempered_transition(self._h, int(N_t), C.c_double(beta_N_t),
                                                              C.byref(logA), C.byref(acc)))
        return logA.value, bool(acc.value)

    def run_mtt(self, n_iter, n_temp_trans, N_t, beta_N_t):
        self._chk(self._lib.bfmmm_sampler_run_mtt(self._h, int(n_iter), int(n_temp_trans), int(N_t),
                                                  C.c_double(beta_N_t)))

    def record(self, directory, r_stored_iters, thinning_num=1):
        d = directory if directory.endswith("/") else directory + "/"
        self._chk(self._lib.bfmmm_sampler_record(self._h, d.encode(), int(r_stored_iters), int(thinning_num)))

    @property
    def batches_written(self):
        return int(self._lib.bfmmm_sampler_batches_written(self._h))

    def profile(self):
        out = np.zeros(3)
        self._chk(self._lib.bfmmm_sampler_profile(self._h, _p(out)))
        return dict(host_s=out[0], wait_s=out[1], push_s=out[2])

    def tt_trace(self, N_t):
        n = 2 * N_t + 1
        ssr, sig = np.zeros(n), np.zeros(n)
        self._chk(self._lib.bfmmm_sampler_tt_trace(self._h, _p(ssr), _p(sig), n))
        return ssr, sig

    @property
    def iteration(self):
        self._lib.bfmmm_sampler_iteration.restype = C.c_int64
        return int(self._lib.bfmmm_sampler_iteration(self._h))

    @property
    def last_accept(self):
        self._lib.bfmmm_sampler_last_accept.restype = C.c_int64
        return int(self._lib.bfmmm_sampler_last_accept(self._h))

    def set_hband(self, Hband):
        hb = np.ascontiguousarray(Hband, dtype=np.float64)
        self._chk(self._lib.bfmmm_sampler_set_hband(self._h, _p(hb)))

    def set_counts(self, sum_half_total, n_points_total):
        self._chk(self._lib.bfmmm_sampler_set_counts(self._h, C.c_double(sum_half_total), C.c_double(n_points_total)))

    def set_allreduce(self, fn):
        """fn(dev_ptr:int, n_doubles:int, stream:int) -> None; called on the engine's stream order."""
        def _cb(ctx, buf, ln, stream):
            try:
                fn(int(buf), int(ln), int(stream or 0))
                return 0
            except Exception as exc:   # pragma: no cover
                print("all-reduce hook failed:", exc)
                return 1
        self._cb = ALLREDUCE_FN(_cb)
        self._chk(self._lib.bfmmm_sampler_set_allreduce(self._h, self._cb, None))

    def enable_nccl(self, rank: int, world: int, exchange_id, libnccl_path: str | None = None):
        """Installs the native NCCL all-reduce (csrc/nccl_hook.cu).  `exchange_id(id_bytes_or_None) -> bytes`
        must return rank 0's 128-byte unique id on every rank (e.g. a torch.distributed broadcast)."""
        if libnccl_path is None:
            try:
                import nvidia.nccl as _n
                import os as _os
                cand = _os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
                libnccl_path = cand if _os.path.exists(cand) else None
            except Exception:
                libnccl_path = None
        path = libnccl_path.encode() if libnccl_path else None
        idb = None
        if rank == 0:
            buf = C.create_string_buffer(128)
            self._chk(self._lib.bfmmm_nccl_unique_id(path, buf))
            idb = buf.raw
        idb = exchange_id(idb)
        comm = C.c_void_p()
        self._chk(self._lib.bfmmm_sampler_enable_nccl(self._h, path, idb, C.c_int(rank), C.c_int(world), C.byref(comm)))
        self._nccl = comm

    def enable_p2p(self, rank: int, world: int, cap: int, allgather):
        """Installs the NVLink peer-memory all-reduce (csrc/p2p_hook.cu).  `allgather(bytes64) -> bytes`
        returns the concatenation of every rank's 64-byte IPC handle in rank order; the caller adds a barrier
        before the first sweep."""
        ctx = C.c_void_p()
        buf = C.create_string_buffer(64)
        self._chk(self._lib.bfmmm_p2p_create(C.c_int(rank), C.c_int(world), C.c_int64(cap), C.byref(ctx), buf))
        handles = allgather(buf.raw)
        assert len(handles) == 64 * world
        self._p2p = ctx
        self._chk(self._lib.bfmmm_sampler_enable_p2p(self._h, ctx, handles))

    # ------------------------------------------------------------------ injected draws + single updates
    def tape(self, values):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
        self._chk(self._lib.bfmmm_sampler_tape(self._h, _p(v), C.c_int64(v.size)))

    @property
    def device_resident(self) -> bool:
        """True when the sweeps run without a host round trip (globals drawn by device kernels)."""
        return bool(self._lib.bfmmm_sampler_device_resident(self._h))

    def set_tick(self, tick: int):
        """positions the random streams of the host_update calls that follow"""
        self._chk(self._lib.bfmmm_sampler_set_tick(self._h, C.c_int64(tick)))

    def tape_left(self):
        self._lib.bfmmm_sampler_tape_left.restype = C.c_int64
        return int(self._lib.bfmmm_sampler_tape_left(self._h))

    def host_update(self, name, *args):
        fn = getattr(self._lib, "bfmmm_host_update_" + name)
        if name in ("pi", "alpha3"):
            slz = _f(args[0])
            self._chk(fn(self._h, _p(slz)))
        elif name in ("phi", "nu", "eta", "xi"):
            WtW, BtYW = _f(args[0]), _f(args[1])
            beta = args[2] if len(args) > 2 else 1.0
            self._chk(fn(self._h, _p(WtW), _p(BtYW), C.c_double(beta)))
        elif name == "sigma":
            ssr, beta, tempered = args
            self._chk(fn(self._h, C.c_double(ssr), C.c_double(beta), int(tempered)))
        else:
            self._chk(fn(self._h))
