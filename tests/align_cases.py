"""Inputs for the RANSAC depth alignment (row f3, second half): relative (scale-invariant) and metric depth maps of a
synthetic scene - metric = scale x relative + noise, a band of gross outliers (a second surface), pixels beyond the
400 m validity limit, infinities in the relative map, a validity mask - at a size the CPU reference finishes in
seconds."""
import numpy as np


def cases():
    out = {}
    for name, (H, W, scale, seed, with_mask, inf_frac) in {
        "room": (60, 80, 2.37, 1, True, 0.02),
        "no_mask": (48, 64, 0.81, 2, False, 0.05),
        "heavy_outliers": (64, 64, 5.5, 3, True, 0.0),
        "tiny": (6, 9, 1.9, 4, True, 0.0),
    }.items():
        rng = np.random.RandomState(seed)
        v, u = np.mgrid[0:H, 0:W]
        rel = (1.0 + 0.8 * u / W + 0.5 * v / H + 0.05 * rng.standard_normal((H, W))).astype(np.float32)
        metric = (scale * rel + 0.02 * scale * rng.standard_normal((H, W))).astype(np.float32)
        frac = 0.35 if name == "heavy_outliers" else 0.1
        bad = rng.random_sample((H, W)) < frac
        metric[bad] = (metric[bad] * rng.uniform(1.5, 3.0, bad.sum())).astype(np.float32)
        far = rng.random_sample((H, W)) < 0.01
        metric[far] = 1000.0
        holes = rng.random_sample((H, W)) < inf_frac if inf_frac else np.zeros((H, W), dtype=bool)
        rel[holes] = np.inf                                # MoGe marks invalid pixels with inf ...
        mask = None
        if with_mask:
            mask = np.ones((H, W), dtype=bool)
            mask[:, : W // 8] = False
            mask[rng.random_sample((H, W)) < 0.05] = False
            mask[holes] = False                            # ... and its mask excludes them
        out[name] = (rel, metric, mask, 100 + seed)
    # an infinite relative depth UNDER the mask: scikit-learn's input check raises out of the reference function
    rel, metric, mask, seed = (a.copy() if hasattr(a, "copy") else a for a in out["room"])
    rel[30, 40] = np.inf
    mask[30, 40] = True
    out["inf_under_mask"] = (rel, metric, mask, seed)
    # nothing valid: the reference returns the metric map
    rel = np.full((4, 5), np.inf, dtype=np.float32)
    out["nothing_valid"] = (rel, np.ones((4, 5), dtype=np.float32), None, 7)
    return out
