"""Stored-sample file layout: our reader parses the reference's own stored
chain files and our writer reproduces them BYTE FOR BYTE (tests/golden/Functional_trace/* are copies
of /root/reference/inst/test-data/Functional_trace/*, written by the reference's warm-start driver,
BFMMM.h:1720-1730)."""
import filecmp
import os

import numpy as np
import pytest

from bayesfmmm_b200 import io as bio

G = os.path.join(os.path.dirname(__file__), "golden")
FT = os.path.join(G, "Functional_trace")


def test_read_reference_text_files():
    sig = bio.load(os.path.join(FT, "Sigma0.txt"))
    assert sig.shape == (150, 1) and abs(np.median(sig[75:]) - 0.00306) < 2e-5
    nu = bio.load(os.path.join(FT, "Nu0.txt"))
    assert nu.shape == (2, 7, 150) and abs(nu[0, 0, 0] - 2.1873619925242878) < 1e-15
    pi = bio.load(os.path.join(FT, "Pi0.txt"))
    assert pi.shape == (2, 150) and np.allclose(pi.sum(axis=0), 1.0)
    a3 = bio.load(os.path.join(FT, "alpha_30.txt"))
    assert a3[0, 0] == 0.0                     # alpha_31(0) is never assigned by the reference
    assert bio.info(os.path.join(FT, "A0.txt")) == (bio.CUBE_TXT, (2, 2, 150, 1, 1))


def test_read_reference_binary_fields():
    kind, dims = bio.info(os.path.join(FT, "Phi0.txt"))
    assert kind == bio.FIELD_CUBE_BIN and dims == (2, 7, 3, 150, 1)
    phi = bio.load(os.path.join(FT, "Phi0.txt"))
    assert phi.shape == (150, 1, 2, 7, 3) and np.all(np.isfinite(phi))
    eta = bio.load(os.path.join(FT, "Eta0.txt"))
    assert eta.shape == (150, 1, 7, 1, 2)
    kind, dims = bio.info(os.path.join(G, "fieldmat.txt"))
    assert kind == bio.FIELD_MAT_BIN


@pytest.mark.parametrize("name", ["Sigma0", "Pi0", "A0", "Nu0", "alpha_30"])
def test_text_writer_is_byte_identical(tmp_path, name):
    src = os.path.join(FT, name + ".txt")
    a = bio.load(src)
    dst = str(tmp_path / (name + ".txt"))
    (bio.save_cube if a.ndim == 3 else bio.save_mat)(dst, a)
    assert filecmp.cmp(src, dst, shallow=False)


@pytest.mark.parametrize("name", ["Phi0", "Eta0"])
def test_binary_field_writer_is_byte_identical(tmp_path, name):
    src = os.path.join(FT, name + ".txt")
    a = bio.load(src)
    dst = str(tmp_path / (name + ".txt"))
    bio.save_field_cube(dst, a)
    assert filecmp.cmp(src, dst, shallow=False)


def test_round_trip_special_values(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(5, 3, 4)) * 10.0 ** rng.integers(-300, 300, size=(5, 3, 4))
    a[0, 0, 0] = 0.0; a[1, 1, 1] = -0.0
    p = str(tmp_path / "c.txt")
    bio.save_cube(p, a)
    assert np.array_equal(bio.load(p), a)          # %.16e round-trips every finite double
    v = rng.normal(size=7)
    bio.save_vec(str(tmp_path / "v.txt"), v)
    assert np.array_equal(bio.load(str(tmp_path / "v.txt")).ravel(), v)
    with pytest.raises(Exception):
        bio.load(str(tmp_path / "missing.txt"))
