"""Launch the run-length path twice on a config-2-shaped batch (for ncu captures): the plain decode
(la3d_rle_decode), then the step as bench.py's rle_input leg runs it (la3d_fit_boxes_rle)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import coco_rle, ops, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
if len(sys.argv) > 2:
    c["B"] = int(sys.argv[2])
B, I, H, W = c["B"], c["I"], c["H"], c["W"]
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + cfg, device="cuda")
host = masks.cpu().numpy().reshape(B * I, H, W)
counts, offsets, max_runs = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in host])
d_counts = torch.as_tensor(counts.view(np.int32), device="cuda")
d_off = torch.as_tensor(offsets, device="cuda")
fitter = ops.RleBoxFitter(B, I, H, W, d_counts.numel(), max_runs, out_dtype=torch.float32)
torch.cuda.synchronize()
for _ in range(2):
    ops.rle_decode(d_counts, d_off, H, W, max_runs)
    fitter(depth, K, d_counts, d_off, ground, "sweep", c["yaw_steps"] or 36, seed=1234)
torch.cuda.synchronize()
print("done", int(d_counts.numel()), "runs, at most", max_runs, "per plane")
