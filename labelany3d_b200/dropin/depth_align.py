"""Drop-in for the depth stage's scale alignment: ``align_depth`` of ``src/batch_scripts/depth.py:52-92``
(same name, signature, prints and return value; that file is a script, so a pipeline imports this function instead
of defining it).

The reference runs scikit-learn's ``RANSACRegressor(LinearRegression(fit_intercept=False), min_samples=0.2)`` on the
valid ``(relative depth, metric depth)`` pixel pairs.  Here the random subsets are drawn exactly as scikit-learn draws
them - ``sample_without_replacement`` on the process-global NumPy generator, one call per trial, the same number of
trials (``_dynamic_max_trials`` is re-evaluated after every better consensus set) - so a seeded run consumes the
generator like the reference does; the arithmetic of a trial (slope of the subset, float32 residuals of all pairs,
inlier count, score and refit sums) and the final map run on the GPU (``csrc/align.cu``), the MAD threshold comes from
the exact float32 median kernel of the same scope-table row.

Parity: the reference's slope is LAPACK's float32 least squares, here it is ``float32(sum(x y) / sum(x x))`` from
float64 sums - the two agree to ~1e-7 relative on the golden cases (tests/golden/make_golden_align.py), not bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch

from labelany3d_b200 import ops as _ops

_EPSILON = np.spacing(1)


def _dynamic_max_trials(n_inliers, n_samples, min_samples, probability):
    inlier_ratio = n_inliers / float(n_samples)
    nom = max(_EPSILON, 1 - probability)
    denom = max(_EPSILON, 1 - inlier_ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def _score(stats):
    n, sy, syy, sres = stats[:4]
    ss_tot = syy - sy * sy / n
    if ss_tot == 0:
        return 1.0 if sres == 0 else 0.0
    return 1.0 - sres / ss_tot


def ransac_slope(x, y, min_samples=0.2, max_trials=100, stop_probability=0.99, random_state=None):
    """The fit of ``RANSACRegressor(LinearRegression(fit_intercept=False), min_samples=...)`` on float32 CUDA vectors
    ``x``, ``y``: returns ``(slope, info)``.  Raises ``ValueError`` with scikit-learn's messages."""
    from sklearn.utils import check_random_state
    from sklearn.utils.random import sample_without_replacement
    n = x.numel()
    m = int(np.ceil(min_samples * n)) if 0 < min_samples < 1 else int(min_samples)
    if m > n:
        raise ValueError("`min_samples` may not be larger than number of samples: n_samples = %d." % n)
    med = _ops.median_f32(y)
    threshold = float(_ops.median_f32((y - med).abs()))
    rng = check_random_state(random_state)
    n_inliers_best, score_best, best = 1, -np.inf, None
    n_trials = 0
    while n_trials < max_trials:
        n_trials += 1
        idx = sample_without_replacement(n, m, random_state=rng)
        coef = _ops.ransac_subset_fit(x, y, torch.as_tensor(idx, dtype=torch.int64, device=x.device))
        stats = _ops.ransac_classify(x, y, coef, threshold)
        n_in = int(stats[0])
        if n_in < n_inliers_best:
            continue
        score = _score(stats)
        if n_in == n_inliers_best and score < score_best:
            continue
        n_inliers_best, score_best, best = n_in, score, stats
        max_trials = min(max_trials, _dynamic_max_trials(n_inliers_best, n, m, stop_probability))
    if best is None:
        raise ValueError("RANSAC could not find a valid consensus set. All `max_trials` iterations were skipped because "
                         "each randomly chosen sub-sample failed the passing criteria. See estimator attributes for "
                         "diagnostics (n_skips*).")
    final = float(np.float32(best[5] / best[4]))               # refit on the inliers of the best trial
    return final, {"n_trials": n_trials, "n_inliers": n_inliers_best, "threshold": threshold, "n": n, "min_samples": m}


def align_depth(relative_depth, metric_depth, mask=None, min_samples=0.2, max_valid_depth=400.0):
    """
    Align scale-invariant depth to metric depth using RANSAC linear regression (``src/batch_scripts/depth.py:52-92``).

    Args:
        relative_depth: Input scale-invariant depth map (e.g., from MoGe), float32.
        metric_depth: Reference metric depth map (e.g., from DepthPro), float32.
        mask: Optional mask to specify valid fitting regions.
        min_samples: Minimum proportion of samples for RANSAC.
        max_valid_depth: Maximum metric depth to be considered valid.

    Returns:
        Aligned metric depth map.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("labelany3d_b200 needs a CUDA device: this path has no CPU implementation")
    dev = torch.device("cuda", torch.cuda.current_device())
    rel_h, met_h = np.asarray(relative_depth), np.asarray(metric_depth)
    if rel_h.dtype != np.float32 or met_h.dtype != np.float32:
        raise TypeError("align_depth: the GPU path takes float32 depth maps (what MoGe / DepthPro produce)")
    rel = torch.as_tensor(np.ascontiguousarray(rel_h), device=dev)
    met = torch.as_tensor(np.ascontiguousarray(met_h), device=dev)
    msk = None if mask is None else torch.as_tensor(np.ascontiguousarray(np.asarray(mask, dtype=bool)), device=dev)
    valid = (~torch.isinf(rel)) & (met < max_valid_depth)
    if msk is not None:
        valid &= msk
    if int(valid.sum()) == 0:
        print("Warning: No valid points for alignment. Returning metric depth.")
        return metric_depth
    try:
        coef, _ = ransac_slope(rel[valid].contiguous(), met[valid].contiguous(), min_samples)
    except Exception as e:  # noqa: BLE001 - the reference catches everything here
        print(f"Error fitting RANSACRegressor: {e}, using metric depth directly")
        return metric_depth
    sel = msk if msk is not None else ~torch.isinf(rel)
    if not bool(torch.isfinite(rel[sel]).all()):            # scikit-learn's input check in predict (:85) raises
        raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
    return _ops.scale_fill(rel, msk, coef).cpu().numpy()
