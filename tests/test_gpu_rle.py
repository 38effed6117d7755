"""GPU parity of the annotation -> mask-stack row through the C ABI and the drop-in:
``la3d_rle_decode`` (COCO run-length planes -> bit planes + quarter counts), ``la3d_fit_boxes_bits`` /
``la3d_fit_boxes_rle`` (the box fit without byte masks) and ``read_bounding_boxes_segmentations``.
Integer work: bit-exact against the oracle (oracle/la3d_oracle_rle.py, pinned against the reference's own
run-length encoder), against the outputs of the unmodified reference loader stored by
tests/golden/make_golden_rle.py, and against ``la3d_mask_scan`` of the decoded byte masks; the boxes are
bit-identical to the ones ``la3d_fit_boxes`` produces from the byte masks."""
import contextlib
import copy
import importlib.util
import io
import json
import os

import numpy as np
import pytest
import torch

import rle_cases
from oracle import la3d_oracle_rle as orr

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


def _dropin(name):
    import labelany3d_b200
    spec = importlib.util.spec_from_file_location(f"_la3d_dropin_{name}", os.path.join(labelany3d_b200.dropin_path(), f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def _packed(run_lists):
    from labelany3d_b200 import coco_rle
    counts, offsets, max_runs = coco_rle.pack_runs([np.asarray(r, dtype=np.uint32) for r in run_lists])
    return dev(counts.view(np.int32)), dev(offsets), max_runs


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def test_rle_decode_matches_the_oracle_and_the_scan(ops):
    cases = rle_cases.codec_masks()
    shapes = sorted({m.shape for _, m in cases})
    for (H, W) in shapes:
        group = [(n, m) for n, m in cases if m.shape == (H, W)]
        runs = [orr.rle_encode_fast(m)["counts"] for _, m in group]
        want_bits, want_cc, want_status = orr.rle_to_bits(runs, H, W)
        counts, offsets, max_runs = _packed(runs)
        for announced in (max_runs, 0, 1, 10 ** 6):          # shared-memory run ends / global workspace / mixed
            bits, cc, status = ops.rle_decode(counts, offsets, H, W, announced)
            np.testing.assert_array_equal(status.cpu().numpy(), want_status, err_msg=f"{H}x{W} max_runs={announced}")
            np.testing.assert_array_equal(_u32(bits), want_bits, err_msg=f"{H}x{W} max_runs={announced}")
            np.testing.assert_array_equal(_u32(cc), want_cc, err_msg=f"{H}x{W} max_runs={announced}")
        # and it is what the mask scan makes of the decoded bytes
        sbits, scc = ops.mask_scan(dev(np.stack([m for _, m in group])))
        assert torch.equal(sbits, bits) and torch.equal(scc, cc)
        np.testing.assert_array_equal(ops.unpack_bits(bits, H, W), np.stack([m for _, m in group]))


def test_rle_decode_edge_cases(ops):
    H, W = 5, 7
    runs = [[0, 40], [3, 0, 0, 2, 0, 5, 1, 1], [2, 3], [0, 0, 0, 35], [], [7], [0, 35], [5, 5, 5, 5, 5, 5, 5], [36],
            [2 ** 32 - 1, 2 ** 32 - 1, 5]]
    want_bits, want_cc, want_status = orr.rle_to_bits(runs, H, W)
    assert want_status.tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 1, 1]
    counts, offsets, max_runs = _packed(runs)
    for announced in (max_runs, 0):
        bits, cc, status = ops.rle_decode(counts, offsets, H, W, announced)
        np.testing.assert_array_equal(status.cpu().numpy(), want_status)
        np.testing.assert_array_equal(_u32(bits), want_bits)
        np.testing.assert_array_equal(_u32(cc), want_cc)
    # a plane with more runs than announced and no workspace is refused (status 2, left empty), the others decode
    from labelany3d_b200 import _lib
    lib = _lib.load()
    chunks, words = ops.scan_layout(H, W)
    P = len(runs)
    bits = torch.full((P, words), -1, dtype=torch.int32, device="cuda")
    cc = torch.full((P, chunks), -1, dtype=torch.int32, device="cuda")
    status = torch.full((P,), -1, dtype=torch.int32, device="cuda")
    rc = lib.la3d_rle_decode(counts.data_ptr(), offsets.data_ptr(), P, H, W, 4, None, bits.data_ptr(), cc.data_ptr(),
                             status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    refused = np.array([len(r) > 4 for r in runs])
    np.testing.assert_array_equal(status.cpu().numpy(), np.where(refused, 2, want_status))
    np.testing.assert_array_equal(_u32(bits)[~refused], want_bits[~refused])
    assert not _u32(bits)[refused].any() and not _u32(cc)[refused].any()


def test_boxes_from_rle_are_the_boxes_from_byte_masks(ops):
    from labelany3d_b200 import synth
    for (B, I, H, W), method, steps in (((3, 4, 120, 160), "pca", 0), ((2, 5, 75, 101), "sweep", 36),
                                        ((2, 3, 96, 128), "convex_hull", 0)):
        depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=77, device="cuda", area=(0.03, 0.2))
        masks[0, 0] = False                                   # an empty plane
        masks[0, 1] = False
        masks[0, 1, 5, 5:9] = True                            # fewer than 500 pixels: no subsample
        want = ops.fit_boxes(depth, K, masks, ground, method, steps, seed=5, image_offset=3)
        host = masks.cpu().numpy().reshape(B * I, H, W)
        runs = [orr.rle_encode_fast(m)["counts"] for m in host]
        counts, offsets, max_runs = _packed(runs)
        fitter = ops.RleBoxFitter(B, I, H, W, counts.numel(), max_runs, device="cuda")
        got = fitter(depth, K, counts, offsets, ground, method, steps, seed=5, image_offset=3)
        torch.cuda.synchronize()
        assert not fitter.rle_status.any()
        assert torch.equal(got.view(torch.int64), want.view(torch.int64)), (method, (got - want).abs().nan_to_num().max())
        bits, cc, _ = ops.rle_decode(counts, offsets, H, W, max_runs)
        again = ops.fit_boxes_bits(depth, K, bits, cc, I, ground, method, steps, seed=5, image_offset=3)
        assert torch.equal(again.view(torch.int64), want.view(torch.int64)), method


def test_full_size_decode_equals_the_scan(ops):
    """BASELINE configs[1] shape (256 x 8 planes of 640x480): the bit planes and quarter counts decoded
    from the runs are the ones the mask scan produces from the byte masks."""
    from labelany3d_b200 import synth
    B, I, H, W = 256, 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + 2, device="cuda")
    host = masks.cpu().numpy().reshape(B * I, H, W)
    runs = [orr.rle_encode_fast(m)["counts"] for m in host]
    counts, offsets, max_runs = _packed(runs)
    bits, cc, status = ops.rle_decode(counts, offsets, H, W, max_runs)
    sbits, scc = ops.mask_scan(masks)
    assert not status.any()
    assert torch.equal(bits, sbits) and torch.equal(cc, scc)
    fitter = ops.RleBoxFitter(B, I, H, W, counts.numel(), max_runs, device="cuda", out_dtype=torch.float32)
    got = fitter(depth, K, counts, offsets, ground, "sweep", 36, seed=1234)
    want = ops.fit_boxes(depth, K, masks, ground, "sweep", 36, seed=1234, out_dtype=torch.float32)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_loader_dropin_matches_the_reference(ops):
    util = _dropin("util")
    annos, size, masks = rle_cases.loader_scene()
    for a in annos:
        seg = a.get("segmentation")
        if isinstance(seg, dict):
            seg["counts"] = orr.rle_to_string(orr.rle_encode_fast(masks[seg.pop("_mask_key")])["counts"]).decode("ascii")
    with open(os.path.join(ROOT, "tests", "golden", "golden_rle_loader_v1.json")) as f:
        want = json.load(f)
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_rle_v1.npz")) as z:
        shape = tuple(z["loader/masks_shape"])
        ref_stack = np.unpackbits(z["loader/masks_packed"], bitorder="little")[:int(np.prod(shape))].reshape(shape).astype(bool)
        ref_ids = z["loader/ids"]
    before = copy.deepcopy(annos)
    with contextlib.redirect_stdout(io.StringIO()) as log:
        bboxes, stack, ids, names = util.read_bounding_boxes_segmentations(annos, size)
    assert annos == before
    assert log.getvalue().splitlines() == want["log"]
    assert bboxes == want["bboxes"] and names == want["names"]
    assert stack.dtype == bool and np.array_equal(stack, ref_stack)
    np.testing.assert_array_equal(ids, ref_ids)
    # uncompressed run lists are accepted too, and a JSON path like the reference
    for a in annos:
        seg = a.get("segmentation")
        if isinstance(seg, dict):
            seg["counts"] = orr.rle_from_string(seg["counts"])
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "annos.json")
        with open(path, "w") as f:
            json.dump(annos, f)
        with contextlib.redirect_stdout(io.StringIO()):
            b2, s2, i2, n2 = util.read_bounding_boxes_segmentations(path, size)
    assert b2 == bboxes and n2 == names and np.array_equal(s2, stack)
    # nothing admitted: np.array([]) and arange(0), like the reference
    with contextlib.redirect_stdout(io.StringIO()):
        b0, s0, i0, n0 = util.read_bounding_boxes_segmentations([], size)
    assert b0 == [] and n0 == [] and s0.shape == (0,) and i0.shape == (0,)
    with pytest.raises(ValueError):
        with contextlib.redirect_stdout(io.StringIO()):
            util.read_bounding_boxes_segmentations([{"iscrowd": 0, "category_id": 1, "bbox": [0, 0, 1, 1],
                                                     "segmentation": {"size": [4, 4], "counts": [0, 17]}}], (4, 4))
    # the polygon / RLE helper of the loader
    W, H = size
    m, h = util.create_boolean_mask_from_polygon(size, [[100.7, 40.2, 190.9, 60.0, 170.3, 190.8, 120.1, 170.5]])
    wm, wh = orr.polygon_mask(size, [[100.7, 40.2, 190.9, 60.0, 170.3, 190.8, 120.1, 170.5]])
    assert np.array_equal(m, wm) and int(h) == int(wh)
    # the helper's RLE branch (dead in the reference's pipeline; it only works for square images there too)
    sq = rle_cases._ellipse(64, 64, 30, 34, 20, 12, 0.5)
    rle = orr.rle_encode_fast(sq)
    m, h = util.create_boolean_mask_from_polygon((64, 64), {"size": rle["size"], "counts": rle["counts"]})
    assert np.array_equal(m, sq) and int(h) == int(np.ptp(np.where(sq.any(1))[0]) + 1)
