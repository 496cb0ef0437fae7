"""ctypes mirror of include/bfmmm_basis.h (B-spline / tensor basis / penalty matrices)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import dp, load_library
from .engine import EngineError

_ip = C.POINTER(C.c_int32)


def _chk(rc):
    if rc != 0:
        raise EngineError(load_library().bfmmm_last_error().decode())


def _p(a):
    return a.ctypes.data_as(dp)


def bspline_basis(t, internal_knots, degree, boundary):
    t = np.ascontiguousarray(t, dtype=np.float64)
    ik = np.ascontiguousarray(internal_knots, dtype=np.float64)
    B = np.zeros((len(t), len(ik) + degree + 1))
    _chk(load_library().bfmmm_bspline_basis(_p(t), C.c_int64(len(t)), _p(ik), len(ik), int(degree),
                                            C.c_double(boundary[0]), C.c_double(boundary[1]), _p(B)))
    return B


def tensor_bspline(t, degrees, boundary, internal_knots):
    """t: n x dim; boundary: dim x 2; internal_knots: one array per dimension (TensorBSpline, BSplines.h:18-62)."""
    t = np.asfortranarray(np.asarray(t, dtype=np.float64))
    n, dim = t.shape
    deg = np.ascontiguousarray(degrees, dtype=np.int32)
    nik = np.ascontiguousarray([len(k) for k in internal_knots], dtype=np.int32)
    ik = np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.float64) for k in internal_knots]))
    bd = np.ascontiguousarray(boundary, dtype=np.float64)
    P = int(np.prod(nik + deg + 1))
    B = np.zeros((n, P))
    _chk(load_library().bfmmm_tensor_bspline(_p(t), C.c_int64(n), dim, deg.ctypes.data_as(_ip), _p(bd), _p(ik),
                                             nik.ctypes.data_as(_ip), _p(B)))
    return B


def get_P(degrees, internal_knots):
    deg = np.ascontiguousarray(degrees, dtype=np.int32)
    nik = np.ascontiguousarray([len(k) for k in internal_knots], dtype=np.int32)
    P = int(np.prod(nik + deg + 1))
    out = np.zeros((P, P), order="F")
    _chk(load_library().bfmmm_get_P(len(deg), deg.ctypes.data_as(_ip), nik.ctypes.data_as(_ip), _p(out)))
    return out


def pmat_rw1(P):
    out = np.zeros((P, P), order="F")
    _chk(load_library().bfmmm_pmat_rw1(int(P), _p(out)))
    return out
