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
