"""Golden vectors for the "next" rows (mask statistics, combine stage, depth-scale alignment),
minted from the UNMODIFIED reference.  Run in the build container:

    python tests/golden/make_golden_next.py

Imports ``src/util.py``, ``src/tools/combine_results.py`` and ``src/model_wrappers.py`` of the
reference in place, feeds them the deterministic inputs of ``tests/next_cases.py``, stores the
reference's outputs in ``tests/golden/golden_next_v1.npz`` / ``golden_combine_v1.json``, checks the
oracle restatement (``oracle/la3d_oracle_next.py``) against them, and writes the Omni3D category
table the combine stage emits (an interface constant of the reference's output format) to
``labelany3d_b200/dropin/coco_omni3d_categories.json``.
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import live_reference  # noqa: E402
import next_cases  # noqa: E402
from oracle import la3d_oracle_next as orn  # noqa: E402

G = {}


def put(k, v):
    G[k] = np.asarray(v)


def main():
    util, _, combine = live_reference.load()
    spec = importlib.util.spec_from_file_location("_la3d_ref_model_wrappers", os.path.join(live_reference.REF_SRC, "model_wrappers.py"))
    mw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mw)

    # ---- f1
    for name, mask, size, bt, st in next_cases.stat_masks():
        trunc, scal = util.analyze_mask(mask, size, scale_threshold=st, boundary_threshold=bt)
        height = util.get_maximum_height(mask)
        put(f"stats/{name}/analyze", [bool(trunc), bool(scal)])
        put(f"stats/{name}/max_height", int(height))
        put(f"stats/{name}/rows", int(np.sum(np.any(mask, axis=1))))       # src/util.py:369-370
        o_tr, o_sc = orn.analyze_mask(mask, size, st, bt)
        assert (bool(o_tr), bool(o_sc)) == (bool(trunc), bool(scal)), name
        assert int(orn.get_maximum_height(mask)) == int(height), name
        s = orn.mask_stats(mask[None], bt)[0]
        assert (s[1] + s[2] + s[3] + s[4] >= 10) == bool(trunc) and (s[0] >= st) == bool(scal), name
        assert (s[6] - s[5] + 1 if s[7] else 0) == int(height), name
    with np.testing.assert_raises(ValueError):
        util.analyze_mask(np.full((4, 4), 2, np.uint8), (4, 4))
    for c, (masks, fg, thr) in enumerate(next_cases.component_cases()):
        a, b = mw.filter_component_masks(masks, fg, thr)
        put(f"components/{c}/fg", a)
        put(f"components/{c}/bg", b)
        oa, ob = orn.filter_component_masks(masks, fg, thr)
        assert np.array_equal(a, oa) and np.array_equal(b, ob)

    # ---- f2
    for c, (a, b) in enumerate(next_cases.iou_box_sets()):
        m = np.array([[combine.iou2D(x, y) for y in b] for x in a])
        put(f"iou/{c}/matrix", m)
        assert np.array_equal(m, orn.iou_matrix(a, b), equal_nan=True)
        if np.isfinite(m).all():
            matches = combine.hungarian_matching(a, b)
            put(f"iou/{c}/matches", np.array([[i, j] for i, j, _ in matches]))
            put(f"iou/{c}/match_iou", np.array([v for _, _, v in matches]))
    with tempfile.TemporaryDirectory() as tmp:
        next_cases.write_results_tree(tmp, "val")
        out = os.path.join(tmp, "COCO3D_val.json")
        log = io.StringIO()
        with contextlib.redirect_stdout(log), contextlib.redirect_stderr(io.StringIO()):
            combine.combine_coco_results(tmp, "val", out)
        with open(out) as f:
            ref_json = json.load(f)
    with open(os.path.join(HERE, "golden_combine_v1.json"), "w") as f:
        json.dump({"output": ref_json, "log": [ln for ln in log.getvalue().splitlines() if ln.startswith(("Warning", "Found", "Saved"))]}, f)
    with open(os.path.join(ROOT, "labelany3d_b200", "dropin", "coco_omni3d_categories.json"), "w") as f:
        json.dump(combine.COCO_CATEGORIES, f, indent=0)
    print(f"combine: {len(ref_json['images'])} images, {len(ref_json['annotations'])} annotations")

    # ---- f3 (align_to_depth_match with a stand-in for the out-of-scope matcher it calls)
    pkg = types.ModuleType("matching")
    sub = types.ModuleType("matching.process_image_space")
    sys.modules["matching"], sys.modules["matching.process_image_space"] = pkg, sub
    for c, case in enumerate(next_cases.align_cases()):
        sub.process_object = lambda object_name, project_root, model, case=case: (case["R"], case["T"], case["render_rgba"], case["depth_render"])
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                T = util.align_to_depth_match(case["mask"], case["depth_map"], "obj", "/nowhere", None)
                n, scale = orn.depth_scale_median(case["mask"], case["depth_map"], case["render_rgba"][..., -1] > 0, case["depth_render"])
        put(f"align/{c}/transform", T)
        put(f"align/{c}/n_overlap", n)
        put(f"align/{c}/scale", np.float32(np.nan) if scale is None else scale)
        if scale is None:
            assert np.array_equal(T, np.eye(4))
        else:
            want = np.eye(4)
            want[:3, :3] = np.linalg.inv(case["R"][:3, :3]) * scale
            want[:3, -1] = case["T"][:3] * scale
            assert np.array_equal(T, want, equal_nan=True), c
    del sys.modules["matching"], sys.modules["matching.process_image_space"]

    np.savez_compressed(os.path.join(HERE, "golden_next_v1.npz"), **G)
    print(f"wrote {len(G)} arrays")


if __name__ == "__main__":
    main()
