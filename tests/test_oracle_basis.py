"""Pins the oracle's basis construction to the reference's two golden files
(src/test-BSplines.cpp:58-82 asserts absdiff <= 1e-7; we require bit equality) and to SciPy."""
import os

import numpy as np

from oracle import oracle as orc
from tests import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "bspline_golden.npz"))


def _tensor_inputs():
    t = np.arange(0, 1000, 10, dtype=float)        # regspace(0,10,990)
    tt = np.stack([t, t], axis=1)
    ik = [np.array([250.0, 500.0, 750.0])] * 2
    return tt, [3, 3], np.array([[0.0, 990.0], [0.0, 990.0]]), ik


def test_tensor_bspline_matches_reference_golden():
    tt, deg, bd, ik = _tensor_inputs()
    B = orc.tensor_bspline(tt, deg, bd, ik)
    assert B.shape == (100, 49)
    assert np.max(np.abs(B - G["tensor_bspline"])) == 0.0


def test_getP_matches_reference_golden():
    _, deg, _, ik = _tensor_inputs()
    Pm = orc.getP(deg, ik)
    assert Pm.shape == (49, 49)
    assert np.max(np.abs(Pm - G["p_mat"])) == 0.0


def test_univariate_basis_matches_scipy_bitwise():
    rng = np.random.default_rng(0)
    t = np.sort(np.concatenate([rng.uniform(0, 1000, 500), [0.0, 1000.0, 250.0, 500.0]]))
    for P, deg in [(7, 3), (8, 3), (20, 3), (6, 2), (5, 1)]:
        ik = synth.equispaced_internal(P, deg)
        B = orc.bspline_basis(t, ik, deg, (0.0, 1000.0))
        Bs = synth.bspline_design(t, ik, deg)
        assert B.shape == (len(t), P)
        assert np.array_equal(B, Bs)
        np.testing.assert_allclose(B.sum(axis=1), 1.0, rtol=0, atol=1e-14)
    # right boundary -> last basis function is exactly 1
    assert B[-1, -1] == 1.0


def test_rw1_penalty():
    Pm = orc.pmat_rw1(5)
    D = np.diff(np.eye(5), axis=0)
    assert np.array_equal(Pm, D.T @ D)
    # univariate GetP equals the tridiagonal P_mat of BFMMM.h:1198-1208
    assert np.array_equal(orc.getP([3], [np.array([250.0, 500.0, 750.0])]), orc.pmat_rw1(7))
