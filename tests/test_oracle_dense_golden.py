"""CPU pin of the all-pixels mode of the oracle (``fit_boxes(..., subsample=False)``) against what the
unmodified reference computed with its random draw replaced by the identity
(tests/golden/make_golden_dense.py; no GPU, no reference tree needed)."""
import os

import numpy as np
import pytest

import dense_cases
from conftest import close
from oracle import la3d_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_dense_v1.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("impl", ["library", "closed"])
def test_all_pixels_hull_oracle_matches_the_reference(gold, impl):
    """method='convex_hull' over ALL masked points (keys records_hull: the unmodified reference with the identity draw)."""
    sel = np.ones(orc.REC, dtype=bool)
    sel[[orc.O_YAW, orc.O_NVALID]] = False
    for name, (depth, K, masks, ground) in dense_cases.scenes().items():
        for use_ground in (0, 1):
            ref = gold[f"{name}/g{use_ground}/records_hull"]
            mine = orc.fit_boxes(depth, K, masks, ground if use_ground else None, "convex_hull", impl=impl, subsample=False)
            np.testing.assert_array_equal(mine[..., orc.O_STATUS], ref[..., orc.O_STATUS])
            for a, b in zip(mine.reshape(-1, orc.REC), ref.reshape(-1, orc.REC)):
                fin = b[sel][np.isfinite(b[sel])]
                scale = max(1.0, float(np.abs(fin).max())) if fin.size else 1.0
                close(a[sel], b[sel], (0.0 if impl == "library" else 1e-9) * scale)


@pytest.mark.parametrize("impl", ["library", "closed"])
def test_all_pixels_oracle_matches_the_reference(gold, impl):
    sel = np.ones(orc.REC, dtype=bool)
    sel[[orc.O_YAW, orc.O_NVALID]] = False
    for name, (depth, K, masks, ground) in dense_cases.scenes().items():
        for use_ground in (0, 1):
            ref = gold[f"{name}/g{use_ground}/records"]
            mine = orc.fit_boxes(depth, K, masks, ground if use_ground else None, "pca", impl=impl, subsample=False)
            np.testing.assert_array_equal(mine[..., orc.O_STATUS], ref[..., orc.O_STATUS])
            np.testing.assert_array_equal(mine[..., orc.O_NMASK], ref[..., orc.O_NMASK])
            for a, b in zip(mine.reshape(-1, orc.REC), ref.reshape(-1, orc.REC)):
                fin = b[sel][np.isfinite(b[sel])]
                scale = max(1.0, float(np.abs(fin).max())) if fin.size else 1.0
                close(a[sel], b[sel], (0.0 if impl == "library" else 1e-9) * scale)
    # masks of at most 500 pixels: the all-pixels mode IS the reference path (no draw happens there)
    depth, K, masks, ground = dense_cases.scenes()["composed"]
    small = orc.mask_counts(masks) <= orc.SUBSAMPLE
    a = orc.fit_boxes(depth, K, masks, ground, "pca", seed=3, impl=impl, subsample=False)
    b = orc.fit_boxes(depth, K, masks, ground, "pca", seed=3, impl=impl, subsample=True)
    assert small.any() and np.array_equal(a[small], b[small], equal_nan=True)
