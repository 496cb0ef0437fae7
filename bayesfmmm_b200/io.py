"""ctypes mirror of include/bfmmm_io.h: the reference's stored-sample files
(ReadVec/ReadMat/ReadCube/ReadFieldCube, src/UserFunctions.cpp:2158-2355)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import dp, load_library
from .engine import EngineError

MAT_TXT, CUBE_TXT, FIELD_CUBE_BIN, MAT_BIN, CUBE_BIN, FIELD_MAT_BIN = 1, 2, 3, 4, 5, 6


def _chk(rc):
    if rc != 0:
        raise EngineError(load_library().bfmmm_last_error().decode())


def _p(a):
    return a.ctypes.data_as(dp)


def save_mat(path, a):
    a = np.asfortranarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))
    if np.asarray(a).ndim == 2 and np.ndim(a) == 2 and a.shape[0] == 1 and np.ndim(np.asarray(a)) == 2:
        pass
    _chk(load_library().bfmmm_save_mat_txt(path.encode(), _p(a), C.c_int64(a.shape[0]), C.c_int64(a.shape[1])))


def save_vec(path, v):
    """Armadillo saves a vec as an n x 1 matrix (Sigma0.txt header '150 1')."""
    save_mat(path, np.asarray(v, dtype=np.float64).reshape(-1, 1))


def save_cube(path, a):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    _chk(load_library().bfmmm_save_cube_txt(path.encode(), _p(a), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]),
                                            C.c_int64(a.shape[2])))


def save_field_cube(path, cubes):
    """cubes: array (n_field_rows, n_field_cols, r, c, s)."""
    cubes = np.asarray(cubes, dtype=np.float64)
    fr, fc, r, c, s = cubes.shape
    flat = np.ascontiguousarray(np.stack([np.asfortranarray(cubes[i, j]).ravel(order="F")
                                          for j in range(fc) for i in range(fr)]))
    _chk(load_library().bfmmm_save_field_cube_bin(path.encode(), _p(flat), C.c_int64(fr), C.c_int64(fc),
                                                  C.c_int64(r), C.c_int64(c), C.c_int64(s)))


def info(path):
    kind = C.c_int32()
    dims = (C.c_int64 * 5)()
    _chk(load_library().bfmmm_file_info(path.encode(), C.byref(kind), dims))
    return kind.value, tuple(int(x) for x in dims)


def load(path):
    """Returns a matrix (r x c), cube (r x c x s) or field array (fr, fc, r, c, s)."""
    kind, (r, c, s, fr, fc) = info(path)
    n = r * c * s * fr * fc
    out = np.zeros(n)
    _chk(load_library().bfmmm_load(path.encode(), _p(out), C.c_int64(n)))
    if kind in (MAT_TXT, MAT_BIN):
        return out.reshape((r, c), order="F")
    if kind in (CUBE_TXT, CUBE_BIN):
        return out.reshape((r, c, s), order="F")
    el = out.reshape((fr * fc, r * c * s))
    arr = np.zeros((fr, fc, r, c, s))
    for j in range(fc):
        for i in range(fr):
            arr[i, j] = el[j * fr + i].reshape((r, c, s), order="F")
    return arr
