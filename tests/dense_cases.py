"""Deterministic inputs of the all-pixels fit tests: shared by tests/golden/make_golden_dense.py (which feeds
them to the unmodified reference) and by the CPU / GPU test-suites."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scenes():
    """``{name: (depth[B,H,W] f32, K[B,3,3], masks[B,I,H,W] bool, ground[B,I,3])}``."""
    from labelany3d_b200 import synth
    from test_oracle_golden import scene_inputs
    out = {}
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz")) as z:
        depth, K, masks, ground, _ = scene_inputs({k: z[k] for k in z.files if k.startswith("scene/")})
    out["composed"] = (depth, K, masks, ground)
    # masks of several thousand pixels; a width that is not a multiple of 32 (bit words run over row ends)
    for name, (B, H, W, I, seed) in {"large": (2, 120, 160, 3, 77), "odd": (2, 75, 101, 3, 78)}.items():
        d, k, m, g = (t.numpy() for t in synth.make_inputs(B, H, W, I, seed=seed, device="cpu", area=(0.1, 0.35)))
        d[0, H // 3, : W // 2] = np.inf           # invalid depths inside masks: rows dropped / box refused
        d[1, H // 2, W // 4: W // 2] = np.nan
        out[name] = (d, k, m, g)
    return out
