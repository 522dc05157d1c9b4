import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 (SURVEY 8(d) parity metric)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def rel_max(a, b):
    """max|a - b| / max|b| (SURVEY 8(d) parity metric)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    mb = np.abs(b).max()
    return np.abs(a - b).max() / (mb if mb > 0 else 1.0)


def assert_parity(out, ref, tol, what=""):
    assert out.shape == ref.shape, f"{what}: shape {out.shape} != {ref.shape}"
    assert np.isfinite(out).all(), f"{what}: non-finite output"
    l2, mx = rel_l2(out, ref), rel_max(out, ref)
    assert l2 <= tol and mx <= tol, f"{what}: rel-L2 {l2:.3e}, rel-max {mx:.3e} > {tol:g}"


@pytest.fixture(scope="session")
def golden():
    return load_golden
