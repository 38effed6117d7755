"""Layout of the packed per-box record (``include/la3d.h``) and helpers."""

from __future__ import annotations

import numpy as np

REC = 64
O_VERT, O_CENTER, O_DIM, O_RCAM = 0, 24, 27, 30
O_YAW, O_NVALID, O_STATUS = 39, 40, 41
O_UV, O_BOX2D, O_NMASK, O_PAD = 42, 58, 62, 63

ST_OK, ST_NO_VALID, ST_PCA_UNDEFINED, ST_BAD_METHOD, ST_NONFINITE, ST_TOO_MANY = 0, 1, 2, 3, 4, 5

FLAG_HULL_FALLBACK = 1      # O_PAD: method convex_hull found no hull and used the PCA yaw (util_3dbox.py:222-224)

METHODS = {"pca": 0, "convex_hull": 1, "sweep": 2}
SUBSAMPLE = 500


def status_error(status, method="pca", n_valid=1):
    """The exception the reference raises where the batch path writes ``status``.

    Messages follow ``src/util_3dbox.py:143,151`` of the reference and, for the two
    cases that surface from scikit-learn inside ``PCA.fit``, scikit-learn's text.
    """
    status = int(status)
    if status == ST_NO_VALID:
        return ValueError("No valid points after removing NaN values")
    if status == ST_BAD_METHOD:
        return ValueError(f"Unknown method: {method}. Use 'pca' or 'convex_hull'")
    if status == ST_NONFINITE:
        return ValueError("Input X contains infinity or a value too large for dtype('float64').")
    if status == ST_PCA_UNDEFINED:
        return ValueError(
            f"n_components=2 must be between 0 and min(n_samples, n_features)={min(int(n_valid), 2)} "
            "with svd_solver='full'")
    if status == ST_TOO_MANY:
        return ValueError("more than 500 points and no sample indices were given")
    return None


def unpack(record):
    """One record (64 scalars) -> the reference's dictionary fields plus the extras."""
    r = np.asarray(record, dtype=np.float64)
    return {
        "bbox3D_cam": r[O_VERT:O_VERT + 24].reshape(8, 3),
        "center_cam": r[O_CENTER:O_CENTER + 3].copy(),
        "dimensions": [r[O_DIM], r[O_DIM + 1], r[O_DIM + 2]],
        "R_cam": r[O_RCAM:O_RCAM + 9].reshape(3, 3),
        "yaw": float(r[O_YAW]),
        "n_valid": int(r[O_NVALID]) if np.isfinite(r[O_NVALID]) else -1,
        "status": int(r[O_STATUS]),
        "corners_2d": r[O_UV:O_UV + 16].reshape(8, 2),
        "bbox2D_proj": r[O_BOX2D:O_BOX2D + 4].copy(),
        "n_mask": int(r[O_NMASK]),
        "hull_fallback": bool(np.isfinite(r[O_PAD]) and int(r[O_PAD]) & FLAG_HULL_FALLBACK),
    }


def bbox2d_trunc(bbox2d_proj, W, H):
    """``bbox2D_trunc`` of ``src/tools/combine_results.py:247-252``."""
    mnx, mny, mxx, mxy = bbox2d_proj
    return [max(0, mnx), max(0, mny), min(W, mxx), min(H, mxy)]
