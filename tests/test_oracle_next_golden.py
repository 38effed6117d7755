"""CPU pins of oracle/la3d_oracle_next.py against the outputs of the unmodified reference stored by
tests/golden/make_golden_next.py (no GPU, no reference tree needed)."""
import json
import os

import numpy as np
import pytest

import next_cases
from oracle import la3d_oracle_next as orn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_next_v1.npz")) as z:
        return {k: z[k] for k in z.files}


def test_mask_statistics(gold):
    for name, mask, size, bt, st in next_cases.stat_masks():
        tr, sc = orn.analyze_mask(mask, size, st, bt)
        assert [bool(tr), bool(sc)] == gold[f"stats/{name}/analyze"].tolist(), name
        assert int(orn.get_maximum_height(mask)) == int(gold[f"stats/{name}/max_height"]), name
        assert int(orn.rows_with_pixels(mask)) == int(gold[f"stats/{name}/rows"]), name
        s = orn.mask_stats(mask[None], bt)[0]
        assert bool(s[1] + s[2] + s[3] + s[4] >= 10) == bool(tr) and bool(s[0] >= st) == bool(sc), name
    with pytest.raises(ValueError):
        orn.analyze_mask(np.full((4, 4), 2, np.uint8), (4, 4))
    for c, (masks, fg, thr) in enumerate(next_cases.component_cases()):
        a, b = orn.filter_component_masks(masks, fg, thr)
        np.testing.assert_array_equal(a, gold[f"components/{c}/fg"])
        np.testing.assert_array_equal(b, gold[f"components/{c}/bg"])


def test_iou_and_matching(gold):
    for c, (a, b) in enumerate(next_cases.iou_box_sets()):
        np.testing.assert_array_equal(orn.iou_matrix(a, b), gold[f"iou/{c}/matrix"])
        if f"iou/{c}/matches" in gold:
            m = orn.hungarian_matching(a, b)
            np.testing.assert_array_equal(np.array([[i, j] for i, j, _ in m]), gold[f"iou/{c}/matches"])
            np.testing.assert_array_equal(np.array([v for _, _, v in m]), gold[f"iou/{c}/match_iou"])


def test_projected_boxes_of_the_combined_json():
    with open(os.path.join(ROOT, "tests", "golden", "golden_combine_v1.json")) as f:
        out = json.load(f)["output"]
    images = {im["id"]: im for im in out["images"]}
    assert len(out["annotations"]) == 31 and len(images) == 7
    for anno in out["annotations"]:
        im = images[anno["image_id"]]
        proj, trunc = orn.box2d_proj_trunc(anno["bbox3D_cam"], np.array(im["K"]), im["width"], im["height"])
        np.testing.assert_allclose(proj, anno["bbox2D_proj"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(np.array(trunc, dtype=float), np.array(anno["bbox2D_trunc"], dtype=float), rtol=1e-12, atol=1e-9)


def test_depth_scale_alignment(gold):
    for c, case in enumerate(next_cases.align_cases()):
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                n, scale = orn.depth_scale_median(case["mask"], case["depth_map"], case["render_rgba"][..., -1] > 0,
                                                  case["depth_render"])
        assert n == int(gold[f"align/{c}/n_overlap"])
        T = gold[f"align/{c}/transform"]
        if scale is None:
            np.testing.assert_array_equal(T, np.eye(4))
        else:
            np.testing.assert_array_equal(scale, gold[f"align/{c}/scale"])
            want = np.eye(4)
            want[:3, :3] = np.linalg.inv(case["R"][:3, :3]) * scale
            want[:3, -1] = case["T"][:3] * scale
            np.testing.assert_array_equal(T, want)
