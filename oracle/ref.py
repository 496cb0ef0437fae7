"""ctypes binding of oracle/_ref/libbfmmm_ref.so: the REFERENCE's own update headers compiled
against oracle/shim/ (see oracle/ref_bridge.cpp).  TEST INFRASTRUCTURE ONLY.

The library can only be (re)built where /root/reference exists (the build container);
on the GPU box the prebuilt .so travels with the snapshot.  `available()` says whether it
is loadable.  Random draws are injected through a FIFO tape laid out in the reference's
call order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libbfmmm_ref.so")
REFERENCE = os.environ.get("BFMMM_REFERENCE", "/root/reference")
_dp = C.POINTER(C.c_double)
_lib = None


def build(force: bool = False) -> bool:
    if not os.path.isdir(os.path.join(REFERENCE, "inst", "include", "BayesFMMM")):
        return os.path.exists(_LIB_PATH)
    src = os.path.join(_HERE, "ref_bridge.cpp")
    shim = os.path.join(_HERE, "shim", "RcppArmadillo.h")
    stale = (not os.path.exists(_LIB_PATH) or
             os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(shim)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"])
    return True


def available() -> bool:
    try:
        return lib() is not None
    except OSError:
        return False


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_tape_left.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O", "F"])


def tape(values):
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
    lib().ref_tape_clear()
    lib().ref_tape_push(_p(v), C.c_int64(v.size))


def tape_left():
    return int(lib().ref_tape_left())


def _done(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed rc={rc}")
    left = tape_left()
    if left:
        raise RuntimeError(f"reference {what}: {left} injected draws were not consumed")


def update_z(d: orc.Data, s: orc.State, pi, alpha3, a_Z_PM, gam, u, beta=1.0, tempered=False):
    """gam: n x K raw gammas, u: n uniforms -> tape order per function: K gammas, 1 uniform."""
    n, K = d.n, d.K
    tape(np.concatenate([np.asarray(gam).reshape(n, K), np.asarray(u).reshape(n, 1)], axis=1))
    out = np.zeros((n, K), order="F")
    dc, sc = d.c(), s.c()
    pi = _f(pi)
    _done(lib().ref_update_z(C.byref(dc), C.byref(sc), _p(pi), C.c_double(alpha3), C.c_double(a_Z_PM),
                             C.c_double(beta), int(tempered), _p(out)), "update_z")
    return out


def update_chi(d, s, eps, beta=1.0, tempered=False):
    tape(np.asarray(eps).reshape(d.n, d.M))          # row-major flatten = (i, m) order
    out = np.zeros((d.n, d.M), order="F")
    dc, sc = d.c(), s.c()
    _done(lib().ref_update_chi(C.byref(dc), C.byref(sc), C.c_double(beta), int(tempered), _p(out)), "update_chi")
    return out


def update_sigma(d, s, alpha0, beta0, gdraw, beta=1.0, tempered=False):
    tape([gdraw])
    out = C.c_double()
    dc, sc = d.c(), s.c()
    _done(lib().ref_update_sigma(C.byref(dc), C.byref(sc), C.c_double(alpha0), C.c_double(beta0),
                                 C.c_double(beta), int(tempered), C.byref(out)), "update_sigma")
    return out.value


def loglik(d, s):
    tape([])
    out = C.c_double()
    dc, sc = d.c(), s.c()
    _done(lib().ref_loglik(C.byref(dc), C.byref(sc), C.byref(out)), "loglik")
    return out.value


def cpo(d, states, burnin_prop=0.0):
    """calcLikelihoodCPO of the reference over the stored iterations `states` (functional models)."""
    tape([])
    dc = d.c()
    scs = [st.c() for st in states]
    arr = (type(scs[0]) * len(scs))(*scs)
    out = np.zeros(d.n)
    _done(lib().ref_cpo(C.byref(dc), arr, len(scs), C.c_double(burnin_prop), out.ctypes.data_as(C.POINTER(C.c_double))), "cpo")
    return out


def update_nu(d, s, tau, Pmat, z, beta=1.0, tempered=False):
    tape(np.asarray(z).ravel(order="F"))             # P x K column-major = block j, then p
    out = np.zeros((d.K, d.P), order="F")
    dc, sc = d.c(), s.c()
    tau = _f(tau); Pm = _f(Pmat) if Pmat is not None else None
    _done(lib().ref_update_nu(C.byref(dc), C.byref(sc), _p(tau), _p(Pm), C.c_double(beta), int(tempered),
                              _p(out)), "update_nu")
    return out


def update_phi(d, s, gamma, tilde_tau, z, beta=1.0, tempered=False):
    tape(np.asarray(z).ravel(order="F"))
    out = np.zeros((d.K, d.P, d.M), order="F")
    dc, sc = d.c(), s.c()
    g = _f(gamma); tt = _f(tilde_tau)
    _done(lib().ref_update_phi(C.byref(dc), C.byref(sc), _p(g), _p(tt), C.c_double(beta), int(tempered),
                               _p(out)), "update_phi")
    return out


def update_eta(d, s, tau_eta, Pmat, z, beta=1.0, tempered=False):
    tape(np.asarray(z).ravel(order="F"))
    out = np.zeros((d.P, d.D, d.K), order="F")
    dc, sc = d.c(), s.c()
    te = _f(tau_eta); Pm = _f(Pmat) if Pmat is not None else None
    _done(lib().ref_update_eta(C.byref(dc), C.byref(sc), _p(te), _p(Pm), C.c_double(beta), int(tempered),
                               _p(out)), "update_eta")
    return out


def update_xi(d, s, gamma_xi, tilde_tau_xi, z, beta=1.0, tempered=False):
    K, P, D, M = d.K, d.P, d.D, d.M
    tape(np.asarray(z).ravel(order="F"))
    g = np.ascontiguousarray(np.stack([np.asfortranarray(gamma_xi[k]).ravel(order="F") for k in range(K)]))
    tt = _f(tilde_tau_xi)
    out = np.zeros((K, P * D * M))
    dc, sc = d.c(), s.c()
    _done(lib().ref_update_xi(C.byref(dc), C.byref(sc), _p(g), _p(tt), C.c_double(beta), int(tempered),
                              _p(out)), "update_xi")
    return np.stack([out[k].reshape((P, D, M), order="F") for k in range(K)])


# ---- host-side prior updates ----
def update_pi(Z, c_hyp, alpha3, a_pi_PM, pi, gam, u):
    n, K = Z.shape
    tape(np.concatenate([np.asarray(gam).ravel(), [u]]))
    Zf, cf, pf = _f(Z), _f(c_hyp), _f(pi)
    out = np.zeros(K)
    _done(lib().ref_update_pi(n, K, _p(Zf), _p(cf), C.c_double(alpha3), C.c_double(a_pi_PM), _p(pf), _p(out)),
          "update_pi")
    return out


def update_alpha3(Z, pi, b, var_alpha3, alpha3, u_prop, u_acc):
    n, K = Z.shape
    tape([u_prop, u_acc])
    Zf, pf = _f(Z), _f(pi)
    out = C.c_double()
    _done(lib().ref_update_alpha3(n, K, _p(Zf), _p(pf), C.c_double(b), C.c_double(var_alpha3),
                                  C.c_double(alpha3), C.byref(out)), "update_alpha3")
    return out.value


def update_tau(nu, Pmat, alpha, beta, gdraws, mv=False):
    K, P = nu.shape
    tape(gdraws)
    nf = _f(nu); Pm = _f(Pmat) if Pmat is not None else None
    out = np.zeros(K)
    _done(lib().ref_update_tau(K, P, _p(nf), _p(Pm), C.c_double(alpha), C.c_double(beta), int(mv), _p(out)),
          "update_tau")
    return out


def update_tau_eta(eta, Pmat, alpha, beta, gdraws, mv=False):
    P, D, K = eta.shape
    tape(gdraws)
    ef = _f(eta); Pm = _f(Pmat) if Pmat is not None else None
    out = np.zeros((K, D), order="F")
    _done(lib().ref_update_tau_eta(K, P, D, _p(ef), _p(Pm), C.c_double(alpha), C.c_double(beta), int(mv),
                                   _p(out)), "update_tau_eta")
    return out


def update_delta(Phi, gamma, A, delta, gdraws):
    K, P, M = Phi.shape
    tape(gdraws)
    pf, gf, af, df = _f(Phi), _f(gamma), _f(A), _f(delta)
    out = np.zeros((K, M), order="F")
    _done(lib().ref_update_delta(K, P, M, _p(pf), _p(gf), _p(af), _p(df), _p(out)), "update_delta")
    return out


def update_gamma(nu_gamma, delta, Phi, gdraws):
    K, P, M = Phi.shape
    tape(gdraws)
    pf, df = _f(Phi), _f(delta)
    out = np.zeros((K, P, M), order="F")
    _done(lib().ref_update_gamma(K, P, M, C.c_double(nu_gamma), _p(df), _p(pf), _p(out)), "update_gamma")
    return out


def update_A(a1l, b1l, a2l, b2l, delta, ve1, ve2, A, us):
    K, M = delta.shape
    tape(us)
    df, af = _f(delta), _f(A)
    out = np.zeros((K, 2), order="F")
    _done(lib().ref_update_A(K, M, C.c_double(a1l), C.c_double(b1l), C.c_double(a2l), C.c_double(b2l), _p(df),
                             C.c_double(ve1), C.c_double(ve2), _p(af), _p(out)), "update_A")
    return out


def _cubes(x, K):
    return np.ascontiguousarray(np.stack([np.asfortranarray(x[k]).ravel(order="F") for k in range(K)]))


def update_delta_xi(xi, gamma_xi, A_xi, delta_xi, gdraws):
    K, P, D, M = xi.shape
    tape(gdraws)
    xf, gf, af, df = _cubes(xi, K), _cubes(gamma_xi, K), _f(A_xi), _f(delta_xi)
    out = np.zeros((K, M, D), order="F")
    _done(lib().ref_update_delta_xi(K, P, M, D, _p(xf), _p(gf), _p(af), _p(df), _p(out)), "update_delta_xi")
    return out


def update_gamma_xi(nu_gamma, delta_xi, xi, gdraws):
    K, P, D, M = xi.shape
    tape(gdraws)
    xf, df = _cubes(xi, K), _f(delta_xi)
    out = np.zeros((K, P * D * M))
    _done(lib().ref_update_gamma_xi(K, P, M, D, C.c_double(nu_gamma), _p(df), _p(xf), _p(out)), "update_gamma_xi")
    return np.stack([out[k].reshape((P, D, M), order="F") for k in range(K)])


def update_A_xi(a1l, b1l, a2l, b2l, delta_xi, ve1, ve2, A_xi, us):
    K, M, D = delta_xi.shape
    tape(us)
    df, af = _f(delta_xi), _f(A_xi)
    out = np.zeros((K, 2, D), order="F")
    _done(lib().ref_update_A_xi(K, M, D, C.c_double(a1l), C.c_double(b1l), C.c_double(a2l), C.c_double(b2l), _p(df),
                                C.c_double(ve1), C.c_double(ve2), _p(af), _p(out)), "update_A_xi")
    return out
