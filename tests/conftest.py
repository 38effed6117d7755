"""pytest configuration: the ``gpu`` marker, import paths, the golden vectors."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Reference outputs minted by tests/golden/make_golden.py (read-only dict)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz")) as z:
        return {k: z[k] for k in z.files}


def opt(a):
    """Golden files store ``None`` as an empty array."""
    return None if a.size == 0 else a


def close(a, b, atol):
    """|a-b| <= atol with NaN==NaN and equal infinities accepted."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    same = (np.isnan(a) & np.isnan(b)) | (np.isinf(a) & (a == b))
    with np.errstate(invalid="ignore"):
        d = np.where(same, 0.0, np.abs(a - b))
    bad = ~(d <= atol)
    assert not bad.any(), f"max|diff|={np.nanmax(d):.3e} at {np.argwhere(bad)[:4].tolist()} (atol {atol})"
