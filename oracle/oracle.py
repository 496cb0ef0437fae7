"""ctypes binding of the CPU oracle (oracle/bfmmm_oracle.cpp).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never from bayesfmmm_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbfmmm_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class _OrcData(C.Structure):
    _fields_ = [("n", C.c_int32), ("K", C.c_int32), ("P", C.c_int32), ("M", C.c_int32),
                ("D", C.c_int32), ("identity_basis", C.c_int32),
                ("off", C.POINTER(C.c_int64)), ("y", _dp), ("B", _dp), ("X", _dp)]


class _OrcState(C.Structure):
    _fields_ = [("nu", _dp), ("Phi", _dp), ("eta", _dp), ("xi", _dp), ("Z", _dp), ("chi", _dp),
                ("sigma_sq", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the restatement with g++ (oracle/Makefile)."""
    src = os.path.join(_HERE, "bfmmm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _f(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O"] + (["F"] if order == "F" else ["C"]))


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


@dataclass
class Data:
    """Observations. Functional: y/B concatenated over functions (B row-major by point);
    multivariate: y is n x P (identity basis)."""
    n: int
    K: int
    P: int
    M: int
    y: np.ndarray
    B: Optional[np.ndarray] = None        # (sum n_i, P) C-contiguous
    off: Optional[np.ndarray] = None      # n+1 int64
    X: Optional[np.ndarray] = None        # n x D
    identity_basis: bool = False
    _keep: list = field(default_factory=list, repr=False)

    @property
    def D(self):
        return 0 if self.X is None else self.X.shape[1]

    def c(self):
        d = _OrcData()
        d.n, d.K, d.P, d.M, d.D = self.n, self.K, self.P, self.M, self.D
        d.identity_basis = 1 if self.identity_basis else 0
        y = _f(self.y, "F")
        self._keep = [y]
        d.y = _p(y)
        if not self.identity_basis:
            off = np.ascontiguousarray(self.off, dtype=np.int64)
            B = _f(self.B, "C")
            self._keep += [off, B]
            d.off = off.ctypes.data_as(C.POINTER(C.c_int64))
            d.B = _p(B)
        if self.X is not None:
            X = _f(self.X, "F")
            self._keep.append(X)
            d.X = _p(X)
        return d


@dataclass
class State:
    nu: np.ndarray            # K x P
    Phi: np.ndarray           # K x P x M
    Z: np.ndarray             # n x K
    chi: np.ndarray           # n x M
    sigma_sq: float
    eta: Optional[np.ndarray] = None     # P x D x K
    xi: Optional[np.ndarray] = None      # K x P x D x M  (cube k = xi[k])
    _keep: list = field(default_factory=list, repr=False)

    def c(self):
        s = _OrcState()
        nu, Phi, Z, chi = _f(self.nu), _f(self.Phi), _f(self.Z), _f(self.chi)
        self._keep = [nu, Phi, Z, chi]
        s.nu, s.Phi, s.Z, s.chi = _p(nu), _p(Phi), _p(Z), _p(chi)
        s.sigma_sq = float(self.sigma_sq)
        if self.eta is not None:
            eta = _f(self.eta)
            self._keep.append(eta)
            s.eta = _p(eta)
        if self.xi is not None:
            # K cubes, each P x D x M column-major, stored back to back
            xi = np.ascontiguousarray(np.stack([np.asfortranarray(self.xi[k]).ravel(order="F")
                                                for k in range(self.xi.shape[0])]))
            self._keep.append(xi)
            s.xi = _p(xi)
        return s


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


# ------------------------------------------------------------------ basis
def bspline_basis(t, internal_knots, degree, boundary):
    t = np.ascontiguousarray(t, dtype=np.float64)
    ik = np.ascontiguousarray(internal_knots, dtype=np.float64)
    P = len(ik) + degree + 1
    B = np.zeros((len(t), P))
    lib().orc_bspline_basis(_p(t), C.c_int64(len(t)), _p(ik), len(ik), degree,
                            C.c_double(boundary[0]), C.c_double(boundary[1]), _p(B))
    return B


def tensor_bspline(t, degrees, boundary, internal_knots):
    """t: n x dim; boundary: dim x 2; internal_knots: list of arrays."""
    t = _f(t, "F")
    n, dim = t.shape
    deg = np.ascontiguousarray(degrees, dtype=np.int32)
    nik = np.ascontiguousarray([len(k) for k in internal_knots], dtype=np.int32)
    ik = np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.float64) for k in internal_knots]))
    bd = np.ascontiguousarray(boundary, dtype=np.float64)
    P = int(np.prod(nik + deg + 1))
    B = np.zeros((n, P))
    lib().orc_tensor_bspline(_p(t), C.c_int64(n), dim, deg.ctypes.data_as(_ip), _p(bd), _p(ik),
                             nik.ctypes.data_as(_ip), _p(B))
    return B


def getP(degrees, internal_knots):
    deg = np.ascontiguousarray(degrees, dtype=np.int32)
    nik = np.ascontiguousarray([len(k) for k in internal_knots], dtype=np.int32)
    P = int(np.prod(nik + deg + 1))
    out = np.zeros((P, P), order="F")
    lib().orc_getP(len(deg), deg.ctypes.data_as(_ip), nik.ctypes.data_as(_ip), _p(out))
    return out


def pmat_rw1(P):
    out = np.zeros((P, P), order="F")
    lib().orc_pmat_rw1(P, _p(out))
    return out


# ------------------------------------------------------------------ updates
def update_z(d: Data, s: State, pi, alpha3, a_Z_PM, gam, u, beta=1.0):
    n, K = d.n, d.K
    pi = _f(pi); gam = _f(gam); u = _f(u)
    Zo = np.zeros((n, K), order="F"); acc = np.zeros(n); took = np.zeros(n, dtype=np.int32)
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_z(C.byref(dc), C.byref(sc), _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM),
                            C.c_double(beta), _p(gam), _p(u), _p(Zo), _p(acc),
                            took.ctypes.data_as(_ip)), "update_z")
    return Zo, acc, took


def update_chi(d: Data, s: State, eps, beta=1.0):
    eps = _f(eps)
    out = np.zeros((d.n, d.M), order="F")
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_chi(C.byref(dc), C.byref(sc), C.c_double(beta), _p(eps), _p(out)), "update_chi")
    return out


def ssr(d: Data, s: State):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    dc, sc = d.c(), s.c()
    _chk(lib().orc_ssr(C.byref(dc), C.byref(sc), C.byref(a), C.byref(b), C.byref(c)), "ssr")
    return a.value, b.value, c.value


def update_sigma(d: Data, s: State, alpha0, beta0, gdraw, beta=1.0, tempered=False):
    sig, a, b = C.c_double(), C.c_double(), C.c_double()
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_sigma(C.byref(dc), C.byref(sc), C.c_double(alpha0), C.c_double(beta0),
                                C.c_double(beta), int(tempered), C.c_double(gdraw), C.byref(sig),
                                C.byref(a), C.byref(b)), "update_sigma")
    return sig.value, a.value, b.value


def loglik(d: Data, s: State):
    ll = C.c_double()
    dc, sc = d.c(), s.c()
    _chk(lib().orc_loglik(C.byref(dc), C.byref(sc), C.byref(ll)), "loglik")
    return ll.value


def marginal_loglik(d: Data, s: State):
    """Per-function marginal log-likelihood (chi integrated out) of one stored iteration: the summand of
    calcLikelihoodCPO (CalculateLikelihood.h:360-375)."""
    out = np.zeros(d.n)
    dc, sc = d.c(), s.c()
    _chk(lib().orc_marginal_loglik(C.byref(dc), C.byref(sc), _p(out)), "marginal_loglik")
    return out


def cpo(logl):
    """CPO_i from the L x n matrix of retained per-iteration marginal log-likelihoods (CalculateLikelihood.h:376-382)."""
    logl = np.asarray(logl, dtype=float)
    mn = logl.min(axis=0)
    return np.log(logl.shape[0]) + mn - np.log(np.exp(mn[None, :] - logl).sum(axis=0))


def update_nu(d: Data, s: State, tau, Pmat, z, beta=1.0):
    tau = _f(tau); z = _f(z)
    Pm = _f(Pmat) if Pmat is not None else None
    out = np.zeros((d.K, d.P), order="F")
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_nu(C.byref(dc), C.byref(sc), _p(tau), _p(Pm), C.c_double(beta), _p(z), _p(out)),
         "update_nu")
    return out


def update_phi(d: Data, s: State, gamma, tilde_tau, z, beta=1.0):
    gamma = _f(gamma); tt = _f(tilde_tau); z = _f(z)
    out = np.zeros((d.K, d.P, d.M), order="F")
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_phi(C.byref(dc), C.byref(sc), _p(gamma), _p(tt), C.c_double(beta), _p(z), _p(out)),
         "update_phi")
    return out


def update_eta(d: Data, s: State, tau_eta, Pmat, z, beta=1.0):
    te = _f(tau_eta); z = _f(z)
    Pm = _f(Pmat) if Pmat is not None else None
    out = np.zeros((d.P, d.D, d.K), order="F")
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_eta(C.byref(dc), C.byref(sc), _p(te), _p(Pm), C.c_double(beta), _p(z), _p(out)),
         "update_eta")
    return out


def update_xi(d: Data, s: State, gamma_xi, tilde_tau_xi, z, beta=1.0):
    """gamma_xi: K x P x D x M; tilde_tau_xi: K x M x D."""
    K, P, D, M = d.K, d.P, d.D, d.M
    g = np.ascontiguousarray(np.stack([np.asfortranarray(gamma_xi[k]).ravel(order="F") for k in range(K)]))
    tt = _f(tilde_tau_xi); z = _f(z)
    out = np.zeros((K, P * D * M))
    dc, sc = d.c(), s.c()
    _chk(lib().orc_update_xi(C.byref(dc), C.byref(sc), _p(g), _p(tt), C.c_double(beta), _p(z), _p(out)),
         "update_xi")
    return np.stack([out[k].reshape((P, D, M), order="F") for k in range(K)])


def pinv_sym(A):
    A = _f(A); out = np.zeros_like(A, order="F")
    _chk(lib().orc_pinv_sym(A.shape[0], _p(A), _p(out)), "pinv")
    return out


def inv(A):
    A = _f(A); out = np.zeros_like(A, order="F")
    _chk(lib().orc_inv(A.shape[0], _p(A), _p(out)), "inv")
    return out
