"""CPU restatement of the depth stage's RANSAC scale alignment.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

Follows ``align_depth``, ``src/batch_scripts/depth.py:52-92`` of the reference, i.e. scikit-learn's
``RANSACRegressor(estimator=LinearRegression(fit_intercept=False), min_samples=0.2).fit`` (pinned
``scikit-learn==1.5.0`` in the reference's ``requirements.txt:23``; restated from 1.9.0's ``_ransac.py``, which is what
runs in this image) with its control flow kept line by line:

* ``min_samples = ceil(0.2 n)``; ``residual_threshold = np.median(|y - np.median(y)|)`` (float32 like the inputs);
* up to 100 trials; each draws ``sample_without_replacement(n, min_samples, random_state)`` from the process-global
  NumPy generator (``check_random_state(None)``), fits the slope through the origin on the subset, classifies
  ``|y - x*coef| <= threshold``, skips the trial if it has fewer inliers than the best so far, scores the inliers
  (R^2), skips on an equal count with a worse score, and re-evaluates ``_dynamic_max_trials``;
* the final model is refitted on the best trial's inliers; the output map is ``coef * relative_depth`` under the mask
  and ``10000.0`` elsewhere.

**Parity status: PINNED to the reference statistically, not bit for bit.**  The reference's slope comes out of LAPACK's
float32 least squares (``scipy.linalg.lstsq`` inside ``LinearRegression``), which is not restated: here (and in the CUDA
kernels) the slope is ``float32(sum(x*y) / sum(x*x))`` with float64 sums.  The draws are the reference's own (same
generator, same calls, same count), so with the same seed the subsets are identical; measured against the live
reference on the cases of ``tests/golden/make_golden_align.py`` the slopes agree to ~1e-7 relative and the generator
ends in the same state.  Exact ties of inlier counts are broken by an R^2 whose total sum of squares is formed in one
float64 pass (scikit-learn centres first); no golden case depends on it.
"""
from __future__ import annotations

import numpy as np
from sklearn.utils import check_random_state
from sklearn.utils.random import sample_without_replacement

_EPSILON = np.spacing(1)
FILL = 10000.0
MSG_NONFINITE = "Input X contains infinity or a value too large for dtype('float32')."


def dynamic_max_trials(n_inliers, n_samples, min_samples, probability):
    inlier_ratio = n_inliers / float(n_samples)
    nom = max(_EPSILON, 1 - probability)
    denom = max(_EPSILON, 1 - inlier_ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def slope(x, y):
    """Least-squares slope through the origin, float64 sums, rounded to the inputs' float32."""
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    return np.float32(np.dot(x64, y64) / np.dot(x64, x64))


def classify(x, y, coef, threshold):
    """``(inlier mask, stats)``: float32 residuals like NumPy's ``y - X @ coef``; stats over the inliers."""
    diff = y - x * np.float32(coef)
    inl = np.abs(diff) <= threshold
    yi, di, xi = y[inl].astype(np.float64), diff[inl].astype(np.float64), x[inl].astype(np.float64)
    stats = np.array([inl.sum(), yi.sum(), (yi * yi).sum(), (di * di).sum(), (xi * xi).sum(), (xi * yi).sum()])
    return inl, stats


def score_from_stats(stats):
    n, sy, syy, sres = stats[:4]
    ss_tot = syy - sy * sy / n
    if ss_tot == 0:
        return 1.0 if sres == 0 else 0.0
    return 1.0 - sres / ss_tot


def ransac_slope(x, y, min_samples=0.2, max_trials=100, stop_probability=0.99, random_state=None,
                 fit=slope, classify_fn=classify):
    """Returns ``(coef float32, info)``; raises ``ValueError`` like scikit-learn when no consensus set is found.
    ``fit`` / ``classify_fn`` let the GPU path plug its kernels into the same loop."""
    n = x.shape[0]
    m = int(np.ceil(min_samples * n)) if 0 < min_samples < 1 else int(min_samples)
    if m > n:
        raise ValueError("`min_samples` may not be larger than number of samples: n_samples = %d." % n)
    med = np.median(y)
    threshold = np.median(np.abs(y - med))
    rng = check_random_state(random_state)
    n_inliers_best, score_best, best = 1, -np.inf, None
    n_trials = 0
    while n_trials < max_trials:
        n_trials += 1
        idx = sample_without_replacement(n, m, random_state=rng)
        coef = fit(x[idx], y[idx]) if fit is slope else fit(idx)
        inl, stats = classify_fn(x, y, coef, threshold)
        n_in = int(stats[0])
        if n_in < n_inliers_best:
            continue
        score = score_from_stats(stats)
        if n_in == n_inliers_best and score < score_best:
            continue
        n_inliers_best, score_best, best = n_in, score, stats
        max_trials = min(max_trials, dynamic_max_trials(n_inliers_best, n, m, stop_probability))
    if best is None:
        raise ValueError("RANSAC could not find a valid consensus set. All `max_trials` iterations were skipped because "
                         "each randomly chosen sub-sample failed the passing criteria. See estimator attributes for "
                         "diagnostics (n_skips*).")
    final = np.float32(best[5] / best[4])                      # refit on the inliers of the best trial
    return final, {"n_trials": n_trials, "n_inliers": n_inliers_best, "threshold": float(threshold), "n": n, "min_samples": m}


def align_depth(relative_depth, metric_depth, mask=None, min_samples=0.2, max_valid_depth=400.0, random_state=None,
                return_info=False):
    """``src/batch_scripts/depth.py:52-92``."""
    valid = (~np.isinf(relative_depth)) & (metric_depth < max_valid_depth)
    if mask is not None:
        valid &= mask
    info = {"coef": None}
    if valid.sum() == 0:
        print("Warning: No valid points for alignment. Returning metric depth.")
        return (metric_depth, info) if return_info else metric_depth
    try:
        coef, info = ransac_slope(relative_depth[valid], metric_depth[valid], min_samples, random_state=random_state)
    except Exception as e:  # noqa: BLE001 - the reference catches everything
        print(f"Error fitting RANSACRegressor: {e}, using metric depth directly")
        return (metric_depth, info) if return_info else metric_depth
    info["coef"] = float(coef)
    depth = np.full_like(relative_depth, FILL)
    sel = mask if mask is not None else ~np.isinf(relative_depth)
    if not np.isfinite(relative_depth[sel]).all():          # scikit-learn's check in predict (:85) raises
        raise ValueError(MSG_NONFINITE)
    depth[sel] = relative_depth[sel] * coef
    return (depth, info) if return_info else depth
