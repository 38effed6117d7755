"""Small run of everything round 2 added or rewrote, for compute-sanitizer (memcheck / racecheck / synccheck):
the sampled fit with a multi-destination sink and the flag protocol, the hull method on nearly collinear footprints,
the all-pixels fit (pca / hull / sweep, with and without the polygon filter), the pipeline split, the lift with its
in-kernel camera, the RANSAC alignment kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
B, I, H, W = 2, 3, 96, 128
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda", area=(0.05, 0.3))
want = ops.fit_boxes(depth, K, masks, ground, "convex_hull", seed=2, out_dtype=torch.float32)
# sink: two destinations + flags, two steps
bufs = [torch.empty((2 * B, I, 64), dtype=torch.float32, device="cuda") for _ in range(2)]
flags = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(2)]
status = torch.zeros(1, dtype=torch.int32).pin_memory()
fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
for e in (1, 2):
    flags[0][1:2].fill_(e)
    sink = _lib.make_sink([b.data_ptr() for b in bufs], False, [f.data_ptr() for f in flags], None, status.data_ptr(), e, 0)
    fitter(depth, K, masks, ground, "convex_hull", seed=2, sink=sink)
torch.cuda.synchronize()
assert torch.equal(bufs[1][:B].view(torch.int32), want.view(torch.int32)) and flags[1][0].item() == 1
# pipeline split
lib.la3d_set_pipeline_images(1)
split = ops.fit_boxes(depth, K, masks, ground, "sweep", 12, seed=2)
lib.la3d_set_pipeline_images(-1)
assert torch.equal(split.view(torch.int64), ops.fit_boxes(depth, K, masks, ground, "sweep", 12, seed=2).view(torch.int64))
# hull on tied / nearly collinear footprints (the round-1 hang)
import tie_cases  # noqa: E402
for name, pc in tie_cases.cases().items():
    ops.fit_points(torch.as_tensor(pc).cuda(), torch.tensor([0, len(pc)]).cuda(), None, None, None, "convex_hull")
# all-pixels fit: small planes (no filter) and planes that need the polygon filter
for method, steps in (("pca", 0), ("convex_hull", 0), ("sweep", 12)):
    ops.fit_boxes_all(depth, K, masks, ground, method=method, yaw_steps=steps)
d2, K2, m2, g2 = synth.make_inputs(1, 240, 320, 2, seed=4, device="cuda", area=(0.2, 0.4))
assert int(m2.view(2, -1).sum(1).min()) > 2048
for method, steps in (("convex_hull", 0), ("sweep", 12)):
    rec = ops.fit_boxes_all(d2, K2, m2, g2, method=method, yaw_steps=steps)
    assert (rec[..., 41] == 0).all()
# depth left in pinned host memory: the address-sorted in-place gather (sort in the footprint arrays); masks with
# fewer samples than threads and an empty one; then the sparse bit-plane stores with a reused, stale workspace
m3 = masks.clone()
m3[0, 0] = 0
m3[0, 1] = 0; m3[0, 1, 40, 30:41] = 1
host_depth = depth.cpu().pin_memory()
for method, steps in (("pca", 0), ("convex_hull", 0), ("sweep", 12)):
    got = ops.fit_boxes(host_depth, K, m3, ground, method, steps, seed=2)
    assert torch.equal(got.view(torch.int64), ops.fit_boxes(depth, K, m3, ground, method, steps, seed=2).view(torch.int64))
plan = ops.BoxFitter(B, I, H, W)
plan.workspace.fill_(0xA5)
for mm in (masks, m3, masks):
    five = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    dense = ops.BoxFitter(B, I, H, W)(depth, K, mm, ground, "sweep", 12, seed=2, events=five).clone()
    assert torch.equal(plan(depth, K, mm, ground, "sweep", 12, seed=2).view(torch.int64), dense.view(torch.int64))
# lift with the in-kernel camera (f32 / f64)
ops.depth_lift(depth, K, out_dtype=torch.float32)
ops.depth_lift(depth, K, out_dtype=torch.float64)
# RANSAC pieces
x = torch.rand(5000, device="cuda") + 1.0
y = 1.7 * x + 0.01 * torch.randn(5000, device="cuda")
idx = torch.randperm(5000, device="cuda")[:1000]
ops.ransac_subset_fit(x, y, idx)
ops.ransac_classify(x, y, 1.7, 0.02)
ops.scale_fill(x.view(50, 100), None, 1.7)
ops.median_f32(y)
torch.cuda.synchronize()
print("sanitize_r2 ok")
