"""Host-side mirror of the C ABI (include/bfmmm.h): one Engine = one shard of functions on one GPU.

Method names follow the reference's update functions (updateZ_PM -> update_z, updateChi ->
update_chi, updateSigma's data pass -> ssr, the accumulations of updateNu/updatePhi/updateEta/
updateXi -> suffstats).  All arrays are NumPy float64 in Armadillo (column-major) layout.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import Config, dp, load_library

FUNCTIONAL, MULTIVARIATE = 0, 1


class EngineError(RuntimeError):
    pass


def _f(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O", "F"])


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


class Engine:
    def __init__(self, *, model: int, n: int, K: int, P: int, M: int, y, B=None, T: int = 0, off=None,
                 t=None, degree: int = 3, internal_knots=None, boundary=(0.0, 1.0), X=None,
                 device: int = 0, global_offset: int = 0, common_grid: bool = True):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.n, self.K, self.P, self.M = n, K, P, M
        self.D = 0 if X is None else np.asarray(X).shape[1]
        self.q = K * (1 + self.D) * (1 + M)
        cfg = Config()
        cfg.model, cfg.n, cfg.K, cfg.P, cfg.M, cfg.D = model, n, K, P, M, self.D
        cfg.device, cfg.common_grid, cfg.T = device, 1 if common_grid else 0, T
        keep = []
        yv = _f(y) if model == MULTIVARIATE else np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        keep.append(yv)
        cfg.y = _p(yv)
        if B is not None:
            Bv = np.ascontiguousarray(B, dtype=np.float64)
            keep.append(Bv)
            cfg.B = _p(Bv)
        if t is not None:
            tv = np.ascontiguousarray(t, dtype=np.float64)
            keep.append(tv)
            cfg.t = _p(tv)
        if internal_knots is not None:
            ik = np.ascontiguousarray(internal_knots, dtype=np.float64)
            keep.append(ik)
            cfg.internal_knots = _p(ik)
            cfg.n_internal = len(ik)
        cfg.degree = degree
        cfg.boundary[0], cfg.boundary[1] = float(boundary[0]), float(boundary[1])
        if off is not None:
            ov = np.ascontiguousarray(off, dtype=np.int64)
            keep.append(ov)
            cfg.off = ov.ctypes.data_as(C.POINTER(C.c_int64))
        if X is not None:
            Xv = _f(X)
            keep.append(Xv)
            cfg.X = _p(Xv)
        cfg.global_offset = global_offset
        rc = self._lib.bfmmm_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise EngineError(self._lib.bfmmm_last_error().decode())

    # ------------------------------------------------------------------ plumbing
    def _chk(self, rc):
        if rc != 0:
            raise EngineError(self._lib.bfmmm_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bfmmm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launch_count(self) -> int:
        return int(self._lib.bfmmm_launch_count())

    @property
    def stream(self) -> int:
        return int(self._lib.bfmmm_stream(self._h) or 0)

    # ------------------------------------------------------------------ state / globals
    def set_state(self, Z=None, chi=None):
        Zf = _f(Z) if Z is not None else None
        cf = _f(chi) if chi is not None else None
        self._chk(self._lib.bfmmm_set_state(self._h, _p(Zf), _p(cf)))

    def get_state(self, Z: bool = True, chi: bool = True):
        Zo = np.zeros((self.n, self.K), order="F") if Z else None
        co = np.zeros((self.n, self.M), order="F") if chi else None
        self._chk(self._lib.bfmmm_get_state(self._h, _p(Zo), _p(co)))
        return Zo, co

    def get_state_rows(self, i0, count, Z: bool = True, chi: bool = True):
        Zo = np.zeros((count, self.K), order="F") if Z else None
        co = np.zeros((count, self.M), order="F") if chi else None
        self._chk(self._lib.bfmmm_get_state_rows(self._h, C.c_int64(i0), C.c_int64(count), _p(Zo), _p(co)))
        return Zo, co

    def get_state_into(self, Z=None, chi=None):
        """D2H copy into caller-owned column-major buffers (the reference's chain slices)."""
        self._chk(self._lib.bfmmm_get_state(self._h, _p(Z), _p(chi)))

    # ------------------------------------------------------------------ post-processing (CPO)
    def marginal_loglik(self):
        """log p(y_i | Z_i, globals, sigma^2), chi_i integrated out (the summand of calcLikelihoodCPO)."""
        out = np.zeros(self.n)
        self._chk(self._lib.bfmmm_marginal_loglik(self._h, _p(out)))
        return out

    def cpo_reset(self):
        self._chk(self._lib.bfmmm_cpo_reset(self._h))

    def cpo_accumulate(self):
        self._chk(self._lib.bfmmm_cpo_accumulate(self._h))

    def cpo_get(self, log_scale=True):
        out = np.zeros(self.n)
        self._chk(self._lib.bfmmm_cpo_get(self._h, _p(out), C.c_int(1 if log_scale else 0)))
        return out

    def get_state_begin(self, Z=None, chi=None):
        """Starts an overlapped read-back of the current (Z, chi) into caller-owned column-major buffers
        (page-locked for the copy to be asynchronous); later updates run while it is in flight."""
        self._chk(self._lib.bfmmm_get_state_begin(self._h, _p(Z), _p(chi)))

    def get_state_wait(self):
        self._chk(self._lib.bfmmm_get_state_wait(self._h))

    def set_globals(self, nu, Phi, sigma_sq, eta=None, xi=None):
        nu, Phi = _f(nu), _f(Phi)
        eta_f = _f(eta) if eta is not None else None
        xi_f = None
        if xi is not None:
            xi_f = np.ascontiguousarray(np.stack([np.asfortranarray(xi[k]).ravel(order="F") for k in range(self.K)]))
        self._chk(self._lib.bfmmm_set_globals(self._h, _p(nu), _p(Phi), _p(eta_f), _p(xi_f), C.c_double(sigma_sq)))

    def seed(self, key: int, iteration: int = 0):
        self._chk(self._lib.bfmmm_seed(self._h, C.c_uint64(key), C.c_uint64(iteration)))

    # ------------------------------------------------------------------ hot path
    def update_z(self, pi, alpha3, a_Z_PM, beta=1.0, gam=None, u=None):
        pi = _f(pi)
        g = _f(gam) if gam is not None else None
        uu = _f(u) if u is not None else None
        slz = np.zeros(self.K)
        nacc = C.c_int64()
        self._chk(self._lib.bfmmm_update_z(self._h, _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM), C.c_double(beta),
                                           _p(g), _p(uu), _p(slz), C.byref(nacc)))
        return slz, nacc.value

    def update_chi(self, beta=1.0, eps=None):
        e = _f(eps) if eps is not None else None
        out = C.c_double()
        self._chk(self._lib.bfmmm_update_chi(self._h, C.c_double(beta), _p(e), C.byref(out)))
        return out.value

    def ssr(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._lib.bfmmm_ssr(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def suffstats(self):
        WtW = np.zeros((self.q, self.q), order="F")
        BtYW = np.zeros((self.P, self.q), order="F")
        self._chk(self._lib.bfmmm_suffstats(self._h, _p(WtW), _p(BtYW)))
        return WtW, BtYW

    def suffstats_ragged(self, bw):
        npairs = self.q * (self.q + 1) // 2
        WtW = np.zeros((self.q, self.q), order="F")
        BtYW = np.zeros((self.P, self.q), order="F")
        Hb = np.zeros((npairs, bw * self.P))
        self._chk(self._lib.bfmmm_suffstats_ragged(self._h, _p(WtW), _p(BtYW), _p(Hb)))
        return WtW, BtYW, Hb

    def dims(self):
        d = (C.c_int32 * 8)()
        self._chk(self._lib.bfmmm_engine_dims(self._h, d))
        return tuple(int(x) for x in d)

    def counts(self):
        """(sum_i floor(n_i / 2), sum_i n_i) of this shard (UpdateSigma.h:49, CalculateLikelihood.h:40)"""
        a, b = C.c_double(), C.c_double()
        self._chk(self._lib.bfmmm_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def gram(self):
        G = np.zeros((self.P, self.P), order="F")
        self._chk(self._lib.bfmmm_get_gram(self._h, _p(G)))
        return G

    def basis(self, T):
        B = np.zeros((T, self.P))
        self._chk(self._lib.bfmmm_get_basis(self._h, _p(B)))
        return B

    # async launches (results stay in the device statistics buffer)
    def update_z_async(self, pi, alpha3, a_Z_PM, beta=1.0):
        pi = _f(pi)
        self._chk(self._lib.bfmmm_update_z_async(self._h, _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM), C.c_double(beta)))

    def update_chi_async(self, beta=1.0):
        self._chk(self._lib.bfmmm_update_chi_async(self._h, C.c_double(beta)))

    def ssr_async(self):
        self._chk(self._lib.bfmmm_ssr_async(self._h))

    def suffstats_async(self):
        self._chk(self._lib.bfmmm_suffstats_async(self._h))

    def sync(self):
        self._chk(self._lib.bfmmm_sync(self._h))

    def stats_buffer(self):
        ptr = dp()
        ln = C.c_int64()
        self._chk(self._lib.bfmmm_stats_buffer_dev(self._h, C.byref(ptr), C.byref(ln)))
        return C.cast(ptr, C.c_void_p).value, ln.value

    def read_stats(self):
        _, ln = self.stats_buffer()
        out = np.zeros(ln)
        self._chk(self._lib.bfmmm_read_stats(self._h, _p(out), C.c_int64(ln)))
        K, q, P = self.K, self.q, self.P
        return dict(sum_log_Z=out[:K], n_accept=out[K], ssr=out[K + 1], ssr_after=out[K + 2],
                    WtW=out[K + 3:K + 3 + q * q].reshape((q, q), order="F"),
                    BtYW=out[K + 3 + q * q:].reshape((P, q), order="F"))

    # ------------------------------------------------------------------ diagnostics (tests)
    def debug_enable_acc(self, on=True):
        self._chk(self._lib.bfmmm_debug_enable_acc(self._h, 1 if on else 0))

    def debug_get_acc(self):
        acc = np.zeros(self.n)
        self._chk(self._lib.bfmmm_debug_get_acc(self._h, _p(acc)))
        return acc

    def debug_update_z_rng(self, pi, alpha3, a_Z_PM, beta=1.0):
        pi = _f(pi)
        gam = np.zeros((self.n, self.K), order="F")
        u = np.zeros(self.n)
        self._chk(self._lib.bfmmm_debug_update_z_rng(self._h, _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM),
                                                     C.c_double(beta), _p(gam), _p(u)))
        return gam, u

    def debug_update_chi_rng(self, beta=1.0):
        eps = np.zeros((self.n, self.M), order="F")
        self._chk(self._lib.bfmmm_debug_update_chi_rng(self._h, C.c_double(beta), _p(eps)))
        return eps

    def debug_z_propose(self, pi, alpha3, a_Z_PM):
        """The proposal half of the Z step on the engine's stream (kernel timing). False when the step is one kernel."""
        pi = np.ascontiguousarray(pi, dtype=np.float64)
        return self._lib.bfmmm_debug_z_propose(self._h, _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM)) == 0

    def debug_moments_valid(self):
        """True when the next update_chi draws from the moments the preceding ssr() left."""
        return bool(self._lib.bfmmm_debug_moments_valid(self._h))

    def debug_get_cache(self):
        Ct = np.zeros((self.n, self.P), order="F")
        rss = np.zeros(self.n)
        self._chk(self._lib.bfmmm_debug_get_cache(self._h, _p(Ct), _p(rss)))
        return Ct, rss
