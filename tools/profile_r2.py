"""Launch the kernels of round 2's bench workloads a few times each (for ncu captures; tools/ncu_round2.sh).

    python tools/profile_r2.py step      # 2048 x 8 x 640x480, 36-step sweep: mask_scan (+prep CTAs), sampler, fit   (x3)
    python tools/profile_r2.py step256   # the same at 256 images (BASELINE configs[1])
    python tools/profile_r2.py lift      # la3d_depth_lift f32 at configs[1] (256 x 640x480) and configs[3] (128 x 1536^2)
    python tools/profile_r2.py fitall    # all-pixels fit (pca, convex_hull, sweep 36) at 256 x 8 x 640x480
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

what = sys.argv[1]
if what in ("step", "step256"):
    B = 2048 if what == "step" else 256
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
    fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
    for _ in range(3):
        fitter(depth, K, masks, ground, "sweep", 36, seed=1234)
elif what == "lift":
    for B, H, W in ((256, 480, 640), (128, 1536, 1536)):
        depth, K, _, _ = synth.make_inputs(2, H, W, 1, seed=1236, device="cuda")
        depth = depth[:1].expand(B, H, W).contiguous()
        K = K[:1].expand(B, 3, 3).contiguous()
        for _ in range(3):
            ops.depth_lift(depth, K, out_dtype=torch.float32)
        del depth, K
        torch.cuda.empty_cache()
elif what == "fitall":
    B, I, H, W = 256, 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
    bits, _ = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I)
    for method, steps in (("pca", 0), ("convex_hull", 0), ("sweep", 36)):
        for _ in range(2):
            ops.fit_all_points(depth, prep, bits, I, torch.float32, method, steps)
torch.cuda.synchronize()
print("done", what)
