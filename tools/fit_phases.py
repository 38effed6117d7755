"""Per-phase latency of the fit kernel's CTAs from clock64 stamps (la3d_debug_fit_clocks): mean cycles between the 8
phase boundaries over all boxes, and the span from the first CTA's start to the last CTA's end on its SM clock."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
names = ["prologue", "gather", "octagon", "yaw", "extents", "record", "stores"]
out = {}
for B in (256, 2048):
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
    bits, cc = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I, seed=1)
    counts, ranks = ops.sample_ranks(cc, B, I, H, W, prep=prep)
    rec = torch.empty((B, I, 64), dtype=torch.float32, device="cuda")
    clocks = torch.zeros((B * I, 8), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for name, mid, steps in (("pca", 0, 0), ("sweep36", 2, 36), ("sweep360", 2, 360), ("hull", 1, 0)):
        for _ in range(3):
            lib.la3d_fit_scanned(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(), ranks.data_ptr(), B, I, H, W, mid, steps, rec.data_ptr(), 0, st)
        lib.la3d_debug_fit_clocks(clocks.data_ptr())
        lib.la3d_fit_scanned(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(), ranks.data_ptr(), B, I, H, W, mid, steps, rec.data_ptr(), 0, st)
        torch.cuda.synchronize()
        lib.la3d_debug_fit_clocks(None)
        c = clocks.double()
        d = (c[:, 1:] - c[:, :-1]).clamp(min=0)
        if mid == 0:
            d[:, 2] = 0
        row = {n: round(float(d[:, i].mean()), 0) for i, n in enumerate(names)}
        row["cta_total"] = round(float((c[:, 7] - c[:, 0]).mean()), 0)
        out[f"B{B}_{name}"] = row
        print(f"B{B}_{name}", row, flush=True)
json.dump(out, open(os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "fit_phases.json"), "w"), indent=1)
