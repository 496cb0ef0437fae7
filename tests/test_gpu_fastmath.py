"""Accuracy of the straight-line FP64 routines of csrc/fastmath.cuh (the Z and chi kernels' logs,
reciprocals, square roots, log-Gammas and Box-Muller normals) against numpy / scipy / mpmath.
The parity bar of the sampler kernels is 1e-10 relative; these routines are held to a few ulp."""
import ctypes as C

import numpy as np
import pytest
from scipy import special, stats

from bayesfmmm_b200._lib import load_library

pytestmark = pytest.mark.gpu
dp = C.POINTER(C.c_double)


def fm(which, x):
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    rc = lib.bfmmm_debug_fastmath(C.c_int(which), x.ctypes.data_as(dp), y.ctypes.data_as(dp), C.c_int64(x.size))
    assert rc == 0, lib.bfmmm_last_error().decode()
    return y


def test_log():
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 200000),
                        1.0 + rng.uniform(-1e-6, 1e-6, 1000), [1.0, 2.0, 0.5, 5e-324, 1e-310, np.inf, 0.0]])
    with np.errstate(divide="ignore"):
        ref = np.log(x)
    got = fm(0, x)
    fin = np.isfinite(ref)
    assert np.array_equal(got[~fin], ref[~fin])
    err = np.abs(got[fin] - ref[fin])
    assert np.all(err <= 3e-16 + 3e-16 * np.abs(ref[fin])), err.max()


def test_rcp_sqrt():
    rng = np.random.default_rng(1)
    x = 10.0 ** rng.uniform(-280, 280, 400000)
    assert np.max(np.abs(fm(1, x) * x - 1.0)) < 4.5e-16
    assert np.max(np.abs(fm(2, x) / np.sqrt(x) - 1.0)) < 2.3e-16


def test_lgamma():
    rng = np.random.default_rng(2)
    x = np.concatenate([10.0 ** rng.uniform(-12, 8, 300000), rng.uniform(0.5, 40.0, 200000), [1.0, 2.0, 16.0, 15.999999]])
    ref = special.gammaln(x)
    got = fm(3, x)
    # absolute floor: below 16 the argument is shifted by 16, so the error is a few ulp of log Gamma(x + 16) ~ 30-80
    bad = np.abs(got - ref) > 6e-14 + 4e-15 * np.abs(ref)
    assert not bad.any(), (x[bad][:5], got[bad][:5], ref[bad][:5])
    import mpmath as mp
    mp.mp.dps = 40
    for xv in [1e-8, 0.3, 1.5, 7.7, 15.9, 16.0, 123.4, 3333.3, 1e7]:
        g = fm(3, np.array([xv]))[0]
        r = float(mp.loggamma(mp.mpf(xv)))
        assert abs(g - r) <= 5e-16 * abs(r) + 5e-14, (xv, g, r)


def test_circle_point():
    rng = np.random.default_rng(3)
    w = rng.integers(0, 2 ** 32, 300000, dtype=np.uint64).astype(np.float64)
    cs, sn = fm(4, w), fm(5, w)
    assert np.max(np.abs(cs * cs + sn * sn - 1.0)) < 5e-16
    wi = w.astype(np.uint64)
    al = ((wi & 0x1FFFFFFF).astype(np.float64) + 0.5) * (np.pi / 4) / 2 ** 29
    s, c = np.sin(al), np.cos(al)
    swap = ((wi >> 29) & 1).astype(bool)
    a, b = np.where(swap, s, c), np.where(swap, c, s)
    a = np.where((wi >> 31) & 1, -a, a)
    b = np.where((wi >> 30) & 1, -b, b)
    assert np.max(np.abs(cs - a)) < 3e-16 and np.max(np.abs(sn - b)) < 3e-16
    # the angle of the point is uniform on the circle
    ang = np.arctan2(sn, cs)
    assert stats.kstest((ang + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3


def test_log1p_series_and_uniform():
    t = np.linspace(-1 / 32, 1 / 32, 100001)
    assert np.max(np.abs(fm(6, t) - np.log1p(t))) < 1e-17 + 2.3e-16 * (1 / 32)
    rng = np.random.default_rng(4)
    b = rng.integers(0, 2 ** 53, 100000, dtype=np.uint64).astype(np.float64)
    u = fm(7, b)
    assert u.min() > 0 and u.max() < 1


def test_box_muller_is_standard_normal():
    idx = np.arange(1, 1_000_001, dtype=np.float64)
    n0, n1 = fm(8, idx), fm(9, idx)
    for z in (n0, n1):
        assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3
        assert stats.kstest(z[::5], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(n0, n1)[0, 1]) < 5e-3
