"""Deterministic inputs for the "next"-row tests (mask statistics, combine stage, depth-scale
alignment).  Shared by tests/golden/make_golden_next.py (which feeds them to the unmodified
reference) and by the CPU / GPU test-suites (which replay them against the stored outputs)."""
import json
import os

import numpy as np


# ------------------------------------------------------------------ f1: masks
def stat_masks():
    """List of (name, mask[H,W] bool, image_size=(W,H), boundary_threshold, scale_threshold)."""
    rng = np.random.RandomState(4242)
    cases = []

    def ellipse(H, W, cy, cx, ry, rx):
        v, u = np.mgrid[:H, :W]
        return ((v - cy) / ry) ** 2 + ((u - cx) / rx) ** 2 <= 1.0

    H, W = 120, 160
    cases.append(("interior", ellipse(H, W, 60, 80, 30, 40), (W, H), 10, 100))
    cases.append(("touch_top", ellipse(H, W, 12, 80, 14, 20), (W, H), 10, 100))
    cases.append(("touch_corner", ellipse(H, W, 5, 5, 12, 12), (W, H), 10, 100))
    cases.append(("touch_right_bottom", ellipse(H, W, 110, 150, 15, 15), (W, H), 10, 100))
    cases.append(("tiny", ellipse(H, W, 60, 80, 3, 3), (W, H), 10, 100))
    cases.append(("empty", np.zeros((H, W), bool), (W, H), 10, 100))
    cases.append(("full", np.ones((H, W), bool), (W, H), 10, 100))
    cases.append(("nine_border_pixels", np.pad(np.ones((1, 9), bool), ((0, H - 1), (20, W - 29))), (W, H), 10, 100))
    cases.append(("band0", ellipse(H, W, 60, 80, 30, 40), (W, H), 0, 100))            # [-0:] is the whole image
    cases.append(("band_gt_size", ellipse(H, W, 60, 80, 30, 40), (W, H), 500, 100))
    cases.append(("band3", ellipse(H, W, 4, 80, 3, 10), (W, H), 3, 20))
    H, W = 75, 101                                                                    # rows straddle the 32-bit words
    cases.append(("odd_interior", ellipse(H, W, 40, 50, 20, 30), (W, H), 10, 100))
    cases.append(("odd_left", ellipse(H, W, 40, 4, 20, 9), (W, H), 10, 100))
    cases.append(("odd_noise", rng.rand(H, W) < 0.02, (W, H), 7, 50))
    cases.append(("two_blobs", ellipse(H, W, 15, 30, 5, 8) | ellipse(H, W, 60, 70, 6, 9), (W, H), 10, 100))
    cases.append(("narrow", ellipse(5, 13, 2, 6, 2, 5), (13, 5), 2, 5))
    return cases


def component_cases():
    """List of (masks[I,H,W] bool, foreground[H,W] bool, threshold)."""
    rng = np.random.RandomState(99)
    out = []
    for (I, H, W, thr) in ((6, 96, 128, 0.5), (5, 75, 101, 0.3), (3, 33, 47, 0.5)):
        fg = np.zeros((H, W), bool)
        fg[H // 4:3 * H // 4, W // 5:4 * W // 5] = True
        masks = np.zeros((I, H, W), bool)
        for i in range(I):
            r0, c0 = rng.randint(0, H - 8), rng.randint(0, W - 8)
            masks[i, r0:r0 + rng.randint(4, H // 2), c0:c0 + rng.randint(4, W // 2)] = True
        masks[I - 1] = False                                   # empty mask: (0 + 1e-6) / (0 + 1e-6) = 1 > thr
        # a mask exactly half inside the foreground (ratio == 0.5 up to the 1e-6 terms)
        masks[0] = False
        masks[0, H // 4 - 2:H // 4 + 2, W // 5:W // 5 + 10] = True
        out.append((masks, fg, thr))
    return out


# ------------------------------------------------------------------ f2: combine stage
CATEGORY_NAMES = ["chair", "dining_table", "car", "person", "potted_plant", "tv", "not_a_coco_thing", "couch"]


def write_results_tree(root, split="val"):
    """A synthetic experimental_results/COCO/<split>/ tree with every branch of combine_coco_results:
    complete scenes, no bboxes.json, missing 3dbbox / cam_params, empty list, unknown category,
    different numbers of 3D and 2D boxes, a box partly behind the image border."""
    rng = np.random.RandomState(777)
    base = os.path.join(root, split)
    os.makedirs(base, exist_ok=True)

    def box(center, dims, yaw):
        l, w, h = dims
        c = np.array([[-l, -w, -h], [l, -w, -h], [l, w, -h], [-l, w, -h], [-l, -w, h], [l, -w, h], [l, w, h], [-l, w, h]]) / 2
        cy, sy = np.cos(yaw), np.sin(yaw)
        R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        return (c @ R.T + np.array(center)), R

    def scene(name, n3d, n2d, with_cam=True, with_3d=True, W=640, H=480, cats=None, shuffle2d=True):
        d = os.path.join(base, name)
        os.makedirs(d, exist_ok=True)
        K = np.array([[0.9 * W, 0, W / 2], [0, 0.9 * W, H / 2], [0, 0, 1.0]])
        if with_cam:
            with open(os.path.join(d, "cam_params.json"), "w") as f:
                json.dump({"K": K.tolist(), "H": H, "W": W}, f)
        annos, tight = [], []
        for j in range(n3d):
            center = [rng.uniform(-1.5, 1.5), rng.uniform(-0.8, 0.8), rng.uniform(2.5, 6.0)]
            dims = rng.uniform(0.3, 1.5, 3)
            corners, R = box(center, dims, rng.uniform(-1, 1))
            cat = (cats or CATEGORY_NAMES)[j % len(cats or CATEGORY_NAMES)]
            annos.append({"obj_id": j, "category_name": cat, "center_cam": center, "R_cam": R.tolist(),
                          "dimensions": [dims[2], dims[1], dims[0]], "bbox3D_cam": corners.tolist()})
            uv = (K @ corners.T).T
            uv = uv[:, :2] / uv[:, 2:]
            tight.append([uv[:, 0].min() + rng.uniform(-4, 4), uv[:, 1].min() + rng.uniform(-4, 4),
                          uv[:, 0].max() + rng.uniform(-4, 4), uv[:, 1].max() + rng.uniform(-4, 4)])
        if with_3d:
            with open(os.path.join(d, "3dbbox.json"), "w") as f:
                json.dump(annos, f)
        if n2d is not None:
            boxes2d = tight[:n2d] + [[rng.uniform(0, 300), rng.uniform(0, 200), rng.uniform(300, 640), rng.uniform(200, 480)]
                                     for _ in range(max(0, n2d - len(tight)))]
            if shuffle2d:
                boxes2d = [boxes2d[i] for i in rng.permutation(len(boxes2d))]
            with open(os.path.join(d, "bboxes.json"), "w") as f:
                json.dump(boxes2d, f)

    scene("000000000139", 5, 5)
    scene("000000000285", 3, None)                       # no bboxes.json: tight = trunc
    scene("000000000632", 4, 6)                          # more 2D than 3D boxes
    scene("000000000724", 6, 3)                          # fewer 2D than 3D boxes
    scene("000000000776", 0, 2)                          # empty 3dbbox.json: skipped
    scene("000000000785", 3, 3, with_cam=False)          # missing cam_params.json
    scene("000000000802", 3, 3, with_3d=False)           # missing 3dbbox.json
    scene("000000000872", 8, 8, W=500, H=375)            # includes the unknown category
    scene("000000000885", 2, 2, cats=["not_a_coco_thing"])   # every box skipped -> no annotations, image kept
    scene("000000001000", 7, 0)                          # empty bboxes.json
    with open(os.path.join(base, "stray_file.txt"), "w") as f:
        f.write("not a scene")
    return base


def iou_box_sets():
    rng = np.random.RandomState(31)
    sets = []
    for n0, n1 in ((4, 4), (1, 7), (6, 2), (3, 3)):
        a = np.sort(rng.uniform(0, 640, (n0, 2, 2)), axis=1).reshape(n0, 4)[:, [0, 2, 1, 3]]
        b = np.sort(rng.uniform(0, 640, (n1, 2, 2)), axis=1).reshape(n1, 4)[:, [0, 2, 1, 3]]
        sets.append((a, b))
    a = np.array([[10.0, 10, 50, 50], [0, 0, 0, 0], [5, 5, 5, 30], [100, 100, 90, 90]])     # identical, degenerate, inverted
    b = np.array([[10.0, 10, 50, 50], [50, 50, 80, 80], [np.nan, 0, 10, 10], [0, 0, np.inf, 10]])
    sets.append((a, b))
    return sets


# ------------------------------------------------------------------ f3: depth-scale alignment
def align_cases():
    """List of dicts: mask[H,W] bool, depth_map[H,W] f32, render_rgba[H,W,4] f32, depth_render[H,W] f32, R[4,4], T[3]."""
    rng = np.random.RandomState(2024)
    out = []

    def rot(ax, ay):
        ca, sa, cb, sb = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay)
        Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
        Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        M = np.eye(4)
        M[:3, :3] = Rx @ Ry
        return M

    def case(H, W, box_m, box_r, kind):
        mask = np.zeros((H, W), bool)
        mask[box_m[0]:box_m[1], box_m[2]:box_m[3]] = True
        alpha = np.zeros((H, W), np.float32)
        alpha[box_r[0]:box_r[1], box_r[2]:box_r[3]] = rng.uniform(0.1, 1.0, (box_r[1] - box_r[0], box_r[3] - box_r[2]))
        rgba = np.concatenate([rng.rand(H, W, 3).astype(np.float32), alpha[..., None]], -1)
        depth_map = rng.uniform(1.5, 6.0, (H, W)).astype(np.float32)
        depth_render = (depth_map / np.float32(rng.uniform(1.5, 3.0)) * rng.uniform(0.9, 1.1, (H, W))).astype(np.float32)
        if kind == "zeros":
            depth_render[::7, ::5] = 0.0                       # division by zero -> inf ratios
        if kind == "nan":
            depth_map[box_m[0] + 1, box_m[2] + 1] = np.nan
        if kind == "ties":
            depth_render = (depth_map / np.float32(2.0)).astype(np.float32)      # every ratio is exactly 2 (or 2 +- 1 ulp)
        if kind == "sentinel":
            depth_map[::3, ::4] = 10000.0
        return dict(mask=mask, depth_map=depth_map, render_rgba=rgba, depth_render=depth_render,
                    R=rot(rng.uniform(-0.5, 0.5), rng.uniform(-1, 1)), T=rng.uniform(-1, 1, 3))

    out.append(case(96, 128, (20, 70, 30, 100), (30, 80, 40, 110), "plain"))       # even / odd counts below
    out.append(case(96, 128, (20, 71, 30, 101), (30, 81, 40, 111), "plain"))
    out.append(case(96, 128, (0, 20, 0, 20), (50, 90, 60, 120), "plain"))          # no overlap
    out.append(case(75, 101, (10, 60, 10, 90), (5, 70, 20, 80), "zeros"))
    out.append(case(75, 101, (10, 60, 10, 90), (5, 70, 20, 80), "nan"))
    out.append(case(75, 101, (10, 60, 10, 90), (5, 70, 20, 80), "ties"))
    out.append(case(120, 160, (10, 110, 10, 150), (0, 120, 0, 160), "sentinel"))
    out.append(case(33, 47, (5, 6, 5, 6), (0, 33, 0, 47), "plain"))                # a single pixel
    out.append(case(33, 47, (5, 6, 5, 7), (0, 33, 0, 47), "plain"))                # two pixels
    return out
