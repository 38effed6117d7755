"""Step time of la3d_fit_boxes as a function of the pipeline split (images per part) and the batch size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
out = {}
for B in (256, 2048):
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
    fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
    for method, steps in (("sweep", 36), ("pca", 0)):
        for per in (0, 32, 64, 128, 256, 512):
            if per >= B and per != 0:
                continue
            lib.la3d_set_pipeline_images(per)
            for _ in range(5):
                fitter(depth, K, masks, ground, method, steps, seed=1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 30
            a.record()
            for _ in range(n):
                fitter(depth, K, masks, ground, method, steps, seed=1)
            b.record()
            torch.cuda.synchronize()
            out[f"B{B}_{method}_per{per}"] = round(a.elapsed_time(b) / n * 1e3, 1)
            print(f"B={B} {method} images/part={per}: {out[f'B{B}_{method}_per{per}']} us", flush=True)
    del depth, K, masks, ground, fitter
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "pipe_sweep.json"), "w"), indent=1)
