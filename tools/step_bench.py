"""Step time of la3d_fit_boxes (one call = scan + sampler + fit) at 256 and 2048 images, sweep-36 and pca: CUDA events
over 30 back-to-back calls.  Environment knobs (LA3D_PDL, LA3D_PIPE_IMAGES, ...) are read by the library."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

out = {}
for B in (256, 2048):
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
    fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
    for method, steps in (("sweep", 36), ("pca", 0)):
        for _ in range(5):
            fitter(depth, K, masks, ground, method, steps, seed=1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30):
            fitter(depth, K, masks, ground, method, steps, seed=1)
        b.record()
        torch.cuda.synchronize()
        out[f"B{B}_{method}{steps or ''}_us"] = round(a.elapsed_time(b) / 30 * 1e3, 1)
    del depth, K, masks, ground, fitter
    torch.cuda.empty_cache()
print(os.environ.get("LA3D_PDL", "0"), out, flush=True)
