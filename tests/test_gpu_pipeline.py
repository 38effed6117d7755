"""The step cut into parts (scan of part p+1 under the sampler + fit of part p on internal high-priority
streams): records must not depend on the split, for byte masks and run-length input, and the call must
remain capturable into a CUDA graph."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def lib():
    from labelany3d_b200 import _lib
    lib = _lib.load()
    yield lib
    lib.la3d_set_pipeline_images(-1)


@pytest.mark.parametrize("method,steps", [("pca", 0), ("sweep", 36), ("convex_hull", 0)])
def test_records_do_not_depend_on_the_split(lib, method, steps):
    from labelany3d_b200 import coco_rle, ops, synth
    B, I, H, W = 11, 3, 96, 160
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=21, device="cuda", area=(0.05, 0.3))
    lib.la3d_set_pipeline_images(0)
    want = ops.fit_boxes(depth, K, masks, ground, method, steps, seed=7)
    counts, offsets, max_runs = coco_rle.runs_from_masks_device(masks.reshape(-1, H, W))
    for per in (1, 2, 4, 5, 11, 64):                      # 11 parts ... 1 part; uneven last parts
        lib.la3d_set_pipeline_images(per)
        for _ in range(3):                                 # back-to-back calls reuse the side streams and events
            got = ops.fit_boxes(depth, K, masks, ground, method, steps, seed=7)
        assert torch.equal(got.view(torch.int64), want.view(torch.int64)), per
        rle = ops.RleBoxFitter(B, I, H, W, counts.numel(), max_runs)(depth, K, counts, offsets, ground, method, steps, seed=7)
        assert torch.equal(rle.view(torch.int64), want.view(torch.int64)), per


def test_split_step_in_a_cuda_graph(lib):
    from labelany3d_b200 import ops, synth
    B, I, H, W = 8, 2, 64, 96
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=22, device="cuda", area=(0.05, 0.3))
    lib.la3d_set_pipeline_images(0)
    want = ops.fit_boxes(depth, K, masks, ground, "sweep", 12, seed=3)
    lib.la3d_set_pipeline_images(3)
    fitter = ops.BoxFitter(B, I, H, W)
    replay, rec = fitter.capture(depth, K, masks, ground, "sweep", 12, seed=3)
    rec.fill_(0)
    replay()
    replay()
    torch.cuda.synchronize()
    assert torch.equal(rec.view(torch.int64), want.view(torch.int64))


def test_headline_batch_equals_its_shards():
    """bench.py's headline workload on one GPU - 2048 images of 640 x 480 with 8 masks in ONE call - gives, bit for bit,
    the records of the eight 256-image shards the 8-GPU run fits (image_offset = the shard's first image), all boxes
    fitted, counts equal to the masks' pixel counts."""
    from labelany3d_b200 import ops, synth
    from oracle import la3d_oracle as orc
    B, I, H, W = 2048, 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
    whole = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)(depth, K, masks, ground, "sweep", 36, seed=1234)
    assert (whole[..., orc.O_STATUS] == 0).all()
    counts = masks.view(B * I, -1).sum(dim=1).view(B, I).float()
    assert torch.equal(whole[..., orc.O_NMASK], counts)
    shard = ops.BoxFitter(256, I, H, W, out_dtype=torch.float32)
    for r in range(8):
        sl = slice(256 * r, 256 * (r + 1))
        part = shard(depth[sl], K[sl], masks[sl], ground[sl], "sweep", 36, seed=1234, image_offset=256 * r)
        assert torch.equal(part.view(torch.int32), whole[sl].contiguous().view(torch.int32)), r
    # a sample against the oracle (the first two images)
    d, k, m, g = (t[:2].cpu().numpy() for t in (depth, K, masks, ground))
    want = orc.fit_boxes(d, k, m, g, "sweep", 36, seed=1234, impl="closed")
    got = whole[:2].cpu().numpy().astype(np.float64)
    ok = np.abs(got[..., :orc.O_YAW] - want[..., :orc.O_YAW]) <= np.maximum(1e-4, 1.2e-7 * np.abs(want[..., :orc.O_YAW]))
    assert ok.all()


@pytest.mark.parametrize("H,W", [(128, 256), (120, 200), (97, 131)])
def test_sparse_bit_planes_from_run_lengths(lib, H, W):
    """The fused run-length step leaves bands of 32 rows without a set pixel unwritten (except the words of chunks
    shared with a neighbouring band): same records as from dense planes, with one plan reused across batches whose
    masks move, for widths whose bands do and do not end on chunk boundaries."""
    from labelany3d_b200 import coco_rle, ops, synth
    B, I = 5, 4
    plan = None
    for seed, area in ((41, (0.3, 0.7)), (42, (0.03, 0.08)), (43, (0.05, 0.3)), (44, (0.03, 0.08))):
        depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=seed, device="cuda", area=area)
        if seed == 42:
            masks[2] = 0
        counts, offsets, max_runs = coco_rle.runs_from_masks_device(masks.reshape(-1, H, W))
        if plan is None:
            plan = ops.RleBoxFitter(B, I, H, W, 4 * counts.numel() + 64, 4 * max_runs + 64)
            plan.workspace.fill_(0xA5)                       # stale bytes where nothing gets written
        bits, cc, status = ops.rle_decode(counts, offsets, H, W, max_runs)          # dense planes
        assert not bool(status.any())
        want = ops.fit_boxes_bits(depth, K, bits, cc, I, ground, "convex_hull", 0, seed=9)
        got = plan(depth, K, counts, offsets, ground, "convex_hull", 0, seed=9)
        assert torch.equal(got.view(torch.int64), want.view(torch.int64)), seed


def test_sparse_bit_planes_and_in_step_events(lib):
    """The fused step does not store the bit words of chunks without set pixels (its workspace is read by the rank
    select only): the records equal those of the step-wise entry points (dense bit planes) bit for bit, also when
    the SAME workspace is reused for masks that empty chunks a previous call filled; and the four events the
    library records around the step's own launches (la3d_debug_step_events) come out ordered."""
    from labelany3d_b200 import ops, synth
    B, I, H, W = 6, 4, 128, 256
    fit = ops.BoxFitter(B, I, H, W)
    fit.workspace.fill_(0xA5)                                # stale bytes where nothing gets written
    for seed, area in ((31, (0.2, 0.6)), (32, (0.02, 0.06)), (33, (0.05, 0.3))):      # large masks first, then small ones
        depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=seed, device="cuda", area=area)
        if seed == 32:
            masks[1] = 0                                                               # a whole image without masks
        five = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        want = ops.BoxFitter(B, I, H, W)(depth, K, masks, ground, "sweep", 36, seed=3, events=five).clone()
        four = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        got = fit(depth, K, masks, ground, "sweep", 36, seed=3, events=four).clone()
        plain = fit(depth, K, masks, ground, "sweep", 36, seed=3).clone()
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int64), want.view(torch.int64)), seed
        assert torch.equal(plain.view(torch.int64), want.view(torch.int64)), seed
        spans = [four[i].elapsed_time(four[i + 1]) for i in range(3)]
        assert all(s > 0 for s in spans), spans
