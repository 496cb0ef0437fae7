"""Product-side basis / penalty construction (include/bfmmm_basis.h) against the reference's golden
files (inst/test-data/Tensor_BSpline.txt, P_mat.txt; src/test-BSplines.cpp:58-82 asserts 1e-7)."""
import os

import numpy as np

from bayesfmmm_b200 import basis
from tests import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "bspline_golden.npz"))


def test_tensor_basis_and_penalty_equal_reference_goldens():
    t = np.arange(0, 1000, 10, dtype=float)
    tt = np.stack([t, t], axis=1)
    ik = [np.array([250.0, 500.0, 750.0])] * 2
    B = basis.tensor_bspline(tt, [3, 3], np.array([[0.0, 990.0], [0.0, 990.0]]), ik)
    assert B.shape == (100, 49) and np.max(np.abs(B - G["tensor_bspline"])) == 0.0
    Pm = basis.get_P([3, 3], ik)
    assert np.max(np.abs(Pm - G["p_mat"])) == 0.0
    assert np.array_equal(basis.get_P([3], [ik[0]]), basis.pmat_rw1(7))


def test_univariate_basis_equals_scipy():
    rng = np.random.default_rng(1)
    t = np.sort(np.concatenate([rng.uniform(0, 1000, 300), [0.0, 1000.0, 500.0]]))
    for P, deg in [(7, 3), (20, 3), (6, 2)]:
        ik = synth.equispaced_internal(P, deg)
        assert np.array_equal(basis.bspline_basis(t, ik, deg, (0.0, 1000.0)), synth.bspline_design(t, ik, deg))
