"""GPU parity of the "next" rows through the C ABI and the drop-ins:
f1 mask statistics, f2 combine stage, f3 depth-scale alignment.  Checked against the outputs of the
unmodified reference stored by tests/golden/make_golden_next.py and, on seeded random inputs, against
the oracle restatement.  Integer work and the float32 median: bit-exact; IoU: bit-exact; projected
2D boxes: 1e-9 relative (np.dot's summation order is BLAS's)."""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import next_cases
from oracle import la3d_oracle_next as orn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__
    __graft_entry__.build()
    from labelany3d_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_next_v1.npz")) as z:
        return {k: z[k] for k in z.files}


def _dropin(name):
    import labelany3d_b200
    spec = importlib.util.spec_from_file_location(f"_la3d_dropin_{name}", os.path.join(labelany3d_b200.dropin_path(), f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


# ------------------------------------------------------------------ f1
def test_mask_stats_golden_and_dropin(ops, gold):
    util = _dropin("util")
    for name, mask, size, bt, st in next_cases.stat_masks():
        H, W = mask.shape
        bits, _ = ops.mask_scan(dev(mask)[None])
        s = ops.mask_stats(bits, H, W, bt).cpu().numpy()[0]
        np.testing.assert_array_equal(s, orn.mask_stats(mask[None], bt)[0], err_msg=name)
        trunc, scal = util.analyze_mask(mask, size, scale_threshold=st, boundary_threshold=bt)
        assert [bool(trunc), bool(scal)] == gold[f"stats/{name}/analyze"].tolist(), name
        assert int(util.get_maximum_height(mask)) == int(gold[f"stats/{name}/max_height"]), name
        assert int(s[ops.STAT_ROWS]) == int(gold[f"stats/{name}/rows"]), name
        # uint8 0/1 masks are binary too
        assert util.analyze_mask(mask.astype(np.uint8), size, st, bt) == (trunc, scal)
    with pytest.raises(ValueError, match="Image Mask must be binary"):
        util.analyze_mask(np.full((4, 4), 2, np.uint8), (4, 4))


@pytest.mark.parametrize("shape", [(3, 5, 96, 128), (2, 3, 75, 101), (1, 2, 480, 640), (1, 1, 7, 13), (2, 2, 37, 384)])
def test_mask_stats_random(ops, shape):
    from labelany3d_b200 import synth
    B, I, H, W = shape
    if H >= 75 and W < 384:
        _, _, masks, _ = synth.make_inputs(B, H, W, I, seed=3, device="cuda", area=(0.01, 0.4))
        masks[0, 0] = False
        masks[0, 0, 0, :] = True
        masks[-1, -1, :, W - 1] = True
    else:
        masks = torch.rand((B, I, H, W), device="cuda") < 0.3
    bits, _ = ops.mask_scan(masks)
    for bt in (10, 1, 0, 33):
        got = ops.mask_stats(bits, H, W, bt).cpu().numpy()
        np.testing.assert_array_equal(got, orn.mask_stats(masks.cpu().numpy(), bt))
    fg = torch.zeros((B, H, W), dtype=torch.bool, device="cuda")
    fg[:, H // 3:, : 2 * W // 3] = True
    fbits, _ = ops.mask_scan(fg)
    inter = ops.mask_overlap(bits, fbits, H, W, group=I).cpu().numpy().reshape(B, I)
    np.testing.assert_array_equal(inter, (masks & fg[:, None]).flatten(2).sum(-1).cpu().numpy())


def test_filter_component_masks_and_admission(ops, gold):
    mf = _dropin("mask_filters")
    for c, (masks, fg, thr) in enumerate(next_cases.component_cases()):
        a, b = mf.filter_component_masks(masks, fg, thr)
        np.testing.assert_array_equal(a, gold[f"components/{c}/fg"])
        np.testing.assert_array_equal(b, gold[f"components/{c}/bg"])
        H, W = masks.shape[-2:]
        keep = mf.admissible_instances(masks, (W, H))
        assert keep.tolist() == [orn.keep_instance(m, (W, H)) for m in masks]


# ------------------------------------------------------------------ f2
def test_iou_matrix_bit_exact(ops, gold):
    cr = _dropin("combine_results")
    for c, (a, b) in enumerate(next_cases.iou_box_sets()):
        got = ops.iou_matrix(dev(a), dev(b)).cpu().numpy()
        np.testing.assert_array_equal(got, gold[f"iou/{c}/matrix"])          # NaN == NaN for assert_array_equal
        if f"iou/{c}/matches" in gold:
            m = cr.hungarian_matching(a, b)
            np.testing.assert_array_equal(np.array([[i, j] for i, j, _ in m]), gold[f"iou/{c}/matches"])
            np.testing.assert_array_equal(np.array([v for _, _, v in m]), gold[f"iou/{c}/match_iou"])
        assert cr.iou2D(a[0], b[0]) == gold[f"iou/{c}/matrix"][0, 0] or np.isnan(gold[f"iou/{c}/matrix"][0, 0])
    # grouped launch = per-group launches
    sets = next_cases.iou_box_sets()[:4]
    off0 = np.cumsum([0] + [len(a) for a, _ in sets])
    off1 = np.cumsum([0] + [len(b) for _, b in sets])
    flat, out_off = ops.iou_matrix(dev(np.concatenate([a for a, _ in sets])), dev(np.concatenate([b for _, b in sets])),
                                   dev(off0.astype(np.int64)), dev(off1.astype(np.int64)))
    flat, out_off = flat.cpu().numpy(), out_off.cpu().numpy()
    for g, (a, b) in enumerate(sets):
        np.testing.assert_array_equal(flat[out_off[g]:out_off[g + 1]].reshape(len(a), len(b)), gold[f"iou/{g}/matrix"])


def _same_json(got, want, path=""):
    assert type(got) is type(want) or (isinstance(got, (int, float)) and isinstance(want, (int, float))), path
    if isinstance(want, dict):
        assert list(got.keys()) == list(want.keys()), path
        for k in want:
            _same_json(got[k], want[k], f"{path}/{k}")
    elif isinstance(want, list):
        assert len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            _same_json(g, w, f"{path}[{i}]")
    elif isinstance(want, float):
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (path, got, want)
    else:
        assert got == want and type(got) is type(want), (path, got, want)


def test_combine_coco_results_equals_the_reference(ops, tmp_path):
    cr = _dropin("combine_results")
    with open(os.path.join(ROOT, "tests", "golden", "golden_combine_v1.json")) as f:
        g = json.load(f)
    next_cases.write_results_tree(str(tmp_path), "val")
    out = tmp_path / "COCO3D_val.json"
    log = io.StringIO()
    with redirect_stdout(log):
        cr.combine_coco_results(str(tmp_path), "val", str(out))
    with open(out) as f:
        got = json.load(f)
    _same_json(got, g["output"])
    lines = [ln for ln in log.getvalue().splitlines() if ln.startswith("Warning")]
    assert lines == [ln for ln in g["log"] if ln.startswith("Warning")]
    with pytest.raises(FileNotFoundError):
        cr.combine_coco_results(str(tmp_path), "train", str(out))


def test_box2d_from_corners_random(ops):
    rng = np.random.RandomState(8)
    n = 300
    corners = rng.uniform(-2, 2, (n, 8, 3)) + np.array([0, 0, 4.0])
    corners[5, 3, 2] = -0.5                       # behind the camera: no special handling in the reference
    K = np.array([[[576.0, 0, 320], [0, 576, 240], [0, 0, 1]], [[450.0, 0, 250], [0, 450, 187.5], [0, 0, 1]]])
    wh = np.array([[640.0, 480], [500, 375]])
    ki = rng.randint(0, 2, n).astype(np.int32)
    proj, trunc = ops.box2d_from_corners(dev(corners), dev(K), dev(wh), dev(ki))
    proj, trunc = proj.cpu().numpy(), trunc.cpu().numpy()
    for i in range(n):
        p, t = orn.box2d_proj_trunc(corners[i], K[ki[i]], wh[ki[i], 0], wh[ki[i], 1])
        np.testing.assert_allclose(proj[i], p, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(trunc[i], np.array(t, dtype=np.float64), rtol=1e-12, atol=1e-9)


# ------------------------------------------------------------------ f3
def test_align_to_depth_match_golden(ops, gold):
    util = _dropin("util")
    pkg, sub = types.ModuleType("matching"), types.ModuleType("matching.process_image_space")
    sys.modules["matching"], sys.modules["matching.process_image_space"] = pkg, sub
    try:
        for c, case in enumerate(next_cases.align_cases()):
            sub.process_object = lambda o, p, m, case=case: (case["R"], case["T"], case["render_rgba"], case["depth_render"])
            with redirect_stdout(io.StringIO()):
                T = util.align_to_depth_match(case["mask"], case["depth_map"], "obj", "/nowhere", None)
            np.testing.assert_array_equal(T, gold[f"align/{c}/transform"], err_msg=str(c))
            H, W = case["mask"].shape
            bits, _ = ops.mask_scan(dev(np.stack([case["mask"], case["render_rgba"][..., -1] > 0])))
            n, scale = ops.depth_scale_median(dev(case["depth_map"])[None], dev(case["depth_render"])[None, None],
                                              bits[:1], bits[1:], H, W)
            assert int(n.item()) == int(gold[f"align/{c}/n_overlap"])
            np.testing.assert_array_equal(scale.cpu().numpy()[0, 0], gold[f"align/{c}/scale"])
    finally:
        del sys.modules["matching"], sys.modules["matching.process_image_space"]


@pytest.mark.parametrize("shape", [(2, 3, 96, 128), (1, 4, 480, 640), (2, 2, 75, 101)])
def test_depth_scale_median_random(ops, shape):
    from labelany3d_b200 import synth
    B, I, H, W = shape
    depth, _, masks, _ = synth.make_inputs(B, H, W, I, seed=21, device="cuda", area=(0.02, 0.3))
    g = torch.Generator(device="cuda").manual_seed(5)
    render = (depth[:, None] / 2.5 * (0.9 + 0.2 * torch.rand((B, I, H, W), device="cuda", generator=g))).contiguous()
    rmask = torch.roll(masks, shifts=(3, -5), dims=(2, 3))
    mb, _ = ops.mask_scan(masks)
    rb, _ = ops.mask_scan(rmask)
    n, scale = ops.depth_scale_median(depth, render, mb, rb, H, W)
    n, scale = n.cpu().numpy(), scale.cpu().numpy()
    d, r, m, rm = depth.cpu().numpy(), render.cpu().numpy(), masks.cpu().numpy(), rmask.cpu().numpy()
    for b in range(B):
        for i in range(I):
            with np.errstate(all="ignore"):
                wn, ws = orn.depth_scale_median(m[b, i], d[b], rm[b, i], r[b, i])
            assert n[b, i] == wn
            if ws is not None:
                assert scale[b, i] == ws and scale.dtype == ws.dtype


# ------------------------------------------------------------------ f4: on-disk formats, the stages chained
def test_scene_directories_round_trip_through_the_pipeline_formats(ops, tmp_path, capsys):
    """depth stage files -> batched fit -> 3dbbox_ground.json -> draw_cube and combine_results consume it."""
    from PIL import Image
    from labelany3d_b200 import scene_io, synth
    from oracle import la3d_oracle as orc
    B, I, H, W = 3, 4, 120, 160
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=9, device="cpu", area=(0.03, 0.2))
    cats = [["chair", "car", "person", "tv"][:I] for _ in range(B)]
    dirs = []
    for b in range(B):
        d = tmp_path / "val" / f"{b:012d}"
        scene_io.save_depth_stage(str(d), depth[b].numpy(), K[b].numpy(), W, H)
        Image.fromarray(np.zeros((H, W, 3), np.uint8)).save(d / "input.png")
        dirs.append(str(d))
    # what the depth stage wrote is what the reference's own consumers expect (src/batch_scripts/depth.py:156-167)
    with open(os.path.join(dirs[0], "cam_params.json")) as f:
        cam = json.load(f)
    assert list(cam.keys()) == ["K", "c2w", "W", "H"] and cam["c2w"] == np.eye(4).tolist()
    assert np.load(os.path.join(dirs[0], "depth_map.npy")).dtype == np.float32
    m = [masks[b].numpy() for b in range(B)]
    m[1] = m[1][:2]                                         # a scene with fewer instances
    cats[1] = cats[1][:2]
    g = [ground[b].numpy()[:len(m[b])] for b in range(B)]
    boxes = scene_io.fit_scene_dirs(dirs, m, cats, g, method="pca", seed=4)
    assert [len(x) for x in boxes] == [4, 2, 4]
    # the same numbers as the oracle's batch path (scene 1 padded with empty masks keeps its own seed)
    want = orc.fit_boxes(depth.numpy(), K.numpy(), masks.numpy(), ground.numpy(), "pca", 0, seed=4, impl="closed")
    for b in (0, 2):
        for i in range(I):
            np.testing.assert_allclose(np.array(boxes[b][i]["bbox3D_cam"]).reshape(-1), want[b, i, :24], rtol=0, atol=1e-8)
    # downstream consumers run on the written files
    util = _dropin("util")
    cr = _dropin("combine_results")
    util.draw_cube(dirs[0], is_ground=True)
    assert os.path.isfile(os.path.join(dirs[0], "vis_3dbox.png"))
    out = tmp_path / "COCO3D_val.json"
    cr.combine_coco_results(str(tmp_path), "val", str(out), bbox_filename="3dbbox_ground.json")
    with open(out) as f:
        combined = json.load(f)
    assert len(combined["images"]) == 3 and len(combined["annotations"]) == 10
    for anno in combined["annotations"]:
        assert anno["bbox2D_tight"] == anno["bbox2D_trunc"] and len(anno["bbox2D_proj"]) == 4
