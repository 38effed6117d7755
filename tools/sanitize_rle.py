"""Small run of the run-length path for compute-sanitizer: both decode paths (W % 128 == 0 and general),
run ends in shared memory and in the global workspace, overflow / empty planes, and the fit from runs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import coco_rle, ops, synth  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


rng = np.random.RandomState(0)
for (H, W) in ((96, 128), (75, 101), (5, 7), (40, 257), (64, 256)):
    masks = [rng.rand(H, W) < d for d in (0.0, 0.02, 0.5, 1.0)]
    v, u = np.mgrid[:H, :W]
    masks.append(((v - H / 2) / (H / 3)) ** 2 + ((u - W / 2) / (W / 4)) ** 2 <= 1.0)
    runs = [coco_rle.runs_from_mask(m) for m in masks] + [np.array([0, H * W + 5], np.uint32), np.zeros(0, np.uint32)]
    counts, offsets, max_runs = coco_rle.pack_runs(runs)
    for announced in (max_runs, 0, 3):
        bits, cc, status = ops.rle_decode(dev(counts.view(np.int32)), dev(offsets), H, W, announced)
        got = ops.unpack_bits(bits, H, W)
        assert all(np.array_equal(got[i], m) for i, m in enumerate(masks)), (H, W, announced)
        assert status.cpu().tolist() == [0] * len(masks) + [1, 0]
B, I, H, W = 2, 3, 96, 128
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda", area=(0.05, 0.2))
host = masks.cpu().numpy().reshape(B * I, H, W)
counts, offsets, max_runs = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in host])
fit = ops.RleBoxFitter(B, I, H, W, counts.size, max_runs)
rec = fit(depth, K, dev(counts.view(np.int32)), dev(offsets), ground, "sweep", 36, seed=1)
want = ops.fit_boxes(depth, K, masks, ground, "sweep", 36, seed=1)
torch.cuda.synchronize()
assert torch.equal(rec.view(torch.int64), want.view(torch.int64))
print("sanitize_rle ok")
