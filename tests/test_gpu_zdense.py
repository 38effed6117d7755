"""GPU parity of the all-pixels fit (``la3d_fit_all_points`` / ``la3d_fit_boxes_all``): the reference's
``estimate_bbox`` (``method='pca'``) with its random 500-point draw replaced by the identity.  Checked against
the records the unmodified reference produced that way (tests/golden/make_golden_dense.py), against the
oracle on seeded inputs, and through properties at BASELINE configs[1] size.  Counts and statuses: exact;
boxes: 1e-9 x scale (float64 records), 1e-4 (float32 records)."""
import os

import numpy as np
import pytest
import torch

import dense_cases
from oracle import la3d_oracle as orc
from test_gpu_parity import TOL_F64, TOL_PRODUCT, check_record

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


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def test_all_pixels_fit_matches_the_reference(ops):
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_dense_v1.npz")) as z:
        gold = {k: z[k] for k in z.files}
    for name, (depth, K, masks, ground) in dense_cases.scenes().items():
        for use_ground in (0, 1):
            ref = gold[f"{name}/g{use_ground}/records"]
            g = dev(ground) if use_ground else None
            rec = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), g).cpu().numpy()
            np.testing.assert_array_equal(rec[..., orc.O_STATUS], ref[..., orc.O_STATUS], err_msg=name)
            np.testing.assert_array_equal(rec[..., orc.O_NMASK], orc.mask_counts(masks), err_msg=name)
            check_record(rec, ref, TOL_F64)
            # what the reference API does not show: surviving points and yaw, against the oracle
            want = orc.fit_boxes(depth, K, masks, ground if use_ground else None, "pca", impl="closed", subsample=False)
            ok = ref[..., orc.O_STATUS] == orc.ST_OK
            np.testing.assert_array_equal(rec[ok][:, orc.O_NVALID], want[ok][:, orc.O_NVALID])
            np.testing.assert_allclose(rec[ok][:, orc.O_YAW], want[ok][:, orc.O_YAW], rtol=0, atol=1e-9)
            bad = ~ok
            np.testing.assert_array_equal(rec[bad][:, orc.O_NVALID], [0.0 if s == 1 else 1.0 for s in ref[bad][:, orc.O_STATUS]])
            # float32 records: the product bar
            rec32 = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), g, out_dtype=torch.float32).cpu().numpy()
            a, b = rec32[ok][:, :orc.O_YAW].astype(np.float64), ref[ok][:, :orc.O_YAW]
            assert (np.abs(a - b) <= np.maximum(TOL_PRODUCT, 1.2e-7 * np.abs(b))).all()


def test_all_pixels_hull_matches_the_reference(ops):
    """method='convex_hull' over every masked pixel against the records the unmodified reference produced with its
    draw replaced by the identity (golden_dense_v1.npz, keys records_hull)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_dense_v1.npz")) as z:
        gold = {k: z[k] for k in z.files}
    for name, (depth, K, masks, ground) in dense_cases.scenes().items():
        for use_ground in (0, 1):
            ref = gold[f"{name}/g{use_ground}/records_hull"]
            g = dev(ground) if use_ground else None
            rec = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), g, method="convex_hull").cpu().numpy()
            np.testing.assert_array_equal(rec[..., orc.O_STATUS], ref[..., orc.O_STATUS], err_msg=name)
            check_record(rec, ref, TOL_F64)


def test_small_masks_equal_the_sampled_path(ops):
    """At most 500 pixels: the reference does not draw, so both kernels must describe the same box."""
    depth, K, masks, ground = dense_cases.scenes()["composed"]
    small = orc.mask_counts(masks) <= orc.SUBSAMPLE
    a = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), dev(ground)).cpu().numpy()
    b = ops.fit_boxes(dev(depth), dev(K), dev(masks), dev(ground), "pca", seed=9).cpu().numpy()
    assert small.any()
    np.testing.assert_array_equal(a[small][:, orc.O_STATUS], b[small][:, orc.O_STATUS])
    check_record(a[small], b[small], TOL_F64, skip=())


def test_full_size_properties(ops):
    """BASELINE configs[1] shape: deterministic (two runs agree bit for bit), independent of how the bit planes
    were made (scan of the bytes / run-length decode), counts exact, a sample of boxes against the oracle."""
    from labelany3d_b200 import coco_rle, synth
    B, I, H, W = 256, 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + 2, device="cuda")
    rec = ops.fit_boxes_all(depth, K, masks, ground)
    again = ops.fit_boxes_all(depth, K, masks, ground)
    assert torch.equal(rec.view(torch.int64), again.view(torch.int64))
    counts = masks.view(B * I, -1).sum(dim=1).view(B, I).double()
    assert torch.equal(rec[..., orc.O_NMASK], counts) and torch.equal(rec[..., orc.O_NVALID], counts)
    assert (rec[..., orc.O_STATUS] == 0).all()
    # from run-length annotations: same bit planes, same boxes
    nb = 16
    host = masks[:nb].cpu().numpy().reshape(nb * I, H, W)
    rc, ro, mr = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in host])
    bits, _, status = ops.rle_decode(dev(rc.view(np.int32)), dev(ro), H, W, mr)
    prep = ops.fit_prepare(K[:nb].contiguous(), ground[:nb].contiguous(), nb, I)
    from_runs = ops.fit_all_points(depth[:nb].contiguous(), prep, bits, I)
    assert not status.any() and torch.equal(from_runs.view(torch.int64), rec[:nb].contiguous().view(torch.int64))
    # a sample against the oracle
    ns = 3
    d, k, m, g = (t[:ns].cpu().numpy() for t in (depth, K, masks, ground))
    want = orc.fit_boxes(d, k, m, g, "pca", impl="closed", subsample=False)
    check_record(rec[:ns].cpu().numpy(), want, TOL_F64, skip=())


@pytest.mark.parametrize("method,steps", [("convex_hull", 0), ("sweep", 36), ("sweep", 360)])
def test_all_pixels_hull_and_sweep_against_the_oracle(ops, method, steps):
    """The yaw searches over EVERY masked pixel: small planes (every point kept, no filter), COCO-size planes
    (the octagon / 16-gon filter leaves at most 2048 candidates) and a 900 x 1100 image whose masks need further
    refinement levels, against the oracle's `subsample=False` restatement (SciPy hull over all points for
    convex_hull)."""
    from labelany3d_b200 import synth
    # composed scene: empty / one-pixel / 200-pixel masks, inf / NaN depths
    depth, K, masks, ground = dense_cases.scenes()["composed"]
    for g in (None, ground):
        want = orc.fit_boxes(depth, K, masks, g, method, steps, impl="closed", subsample=False)
        got = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), None if g is None else dev(g), method=method, yaw_steps=steps).cpu().numpy()
        np.testing.assert_array_equal(got[..., orc.O_STATUS], want[..., orc.O_STATUS])
        ok = want[..., orc.O_STATUS] == orc.ST_OK
        np.testing.assert_array_equal(got[ok][:, orc.O_NVALID], want[ok][:, orc.O_NVALID])
        check_record(got[ok], want[ok], TOL_F64, skip=())
    for (B, I, H, W, area, seed) in ((2, 4, 480, 640, (0.02, 0.12), 31), (1, 3, 900, 1100, (0.05, 0.3), 32)):
        d, k, m, g = synth.make_inputs(B, H, W, I, seed=seed, device="cuda", area=area)
        got = ops.fit_boxes_all(d, k, m, g, method=method, yaw_steps=steps)
        again = ops.fit_boxes_all(d, k, m, g, method=method, yaw_steps=steps)
        assert torch.equal(got.view(torch.int64), again.view(torch.int64))          # deterministic
        want = orc.fit_boxes(d.cpu().numpy(), k.cpu().numpy(), m.cpu().numpy(), g.cpu().numpy(), method, steps,
                             impl="closed", subsample=False)
        assert (want[..., orc.O_NMASK] > 2048).all()
        got = got.cpu().numpy()
        np.testing.assert_array_equal(got[..., orc.O_STATUS], want[..., orc.O_STATUS])
        np.testing.assert_array_equal(got[..., orc.O_NVALID], want[..., orc.O_NVALID])
        np.testing.assert_allclose(got[..., orc.O_YAW], want[..., orc.O_YAW], rtol=0, atol=1e-9)
        check_record(got, want, TOL_F64, skip=())


def test_all_pixels_search_equals_the_sampled_path_on_small_masks(ops):
    depth, K, masks, ground = dense_cases.scenes()["composed"]
    small = orc.mask_counts(masks) <= orc.SUBSAMPLE
    for method, steps in (("convex_hull", 0), ("sweep", 24)):
        a = ops.fit_boxes_all(dev(depth), dev(K), dev(masks), dev(ground), method=method, yaw_steps=steps).cpu().numpy()
        b = ops.fit_boxes(dev(depth), dev(K), dev(masks), dev(ground), method, steps, seed=9).cpu().numpy()
        np.testing.assert_array_equal(a[small][:, orc.O_STATUS], b[small][:, orc.O_STATUS])
        check_record(a[small], b[small], TOL_F64, skip=())


def test_all_pixels_unknown_method_and_rim_overflow(ops):
    from labelany3d_b200 import synth
    d, k, m, g = synth.make_inputs(1, 96, 128, 2, seed=5, device="cuda", area=(0.05, 0.3))
    rec = ops.fit_boxes_all(d, k, m, g, method="nonsense")
    assert (rec[..., orc.O_STATUS] == orc.ST_BAD_METHOD).all()
    # a thin ring: more than 2048 of its points are hull vertices -> status 5 (documented limit), never a wrong box
    H = W = 1024
    v, u = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    r2 = (u - 512.0) ** 2 + (v - 512.0) ** 2
    ring = ((r2 <= 500.0 ** 2) & (r2 >= 499.0 ** 2))[None, None]
    depth = torch.full((1, H, W), 3.0, device="cuda")
    K = torch.tensor([[[900.0, 0, 512], [0, 900.0, 512], [0, 0, 1]]], dtype=torch.float64, device="cuda")
    rec = ops.fit_boxes_all(depth, K, ring, None, method="convex_hull")
    n = int(ring.sum())
    assert n > 2048 and rec[0, 0, orc.O_NMASK] == n
    assert rec[0, 0, orc.O_STATUS] in (0.0, 5.0)
    if rec[0, 0, orc.O_STATUS] == 0.0:                       # fewer than 2048 strict hull vertices after all: then it is right
        want = orc.fit_boxes(depth.cpu().numpy(), K.cpu().numpy(), ring.cpu().numpy(), None, "convex_hull", impl="closed", subsample=False)
        check_record(rec.cpu().numpy(), want, TOL_F64, skip=())
