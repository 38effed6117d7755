"""Time the all-pixels fit kernel alone (bit planes resident) for its three yaw methods on config-2- and
config-4-shaped batches (CUDA events, median of 10).  argv: output JSON name under gpurun_out/."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

out = {}
for tag, (B, I, H, W) in {"cfg2_256x8_640x480": (256, 8, 480, 640), "cfg4_16x20_1536x1536": (16, 20, 1536, 1536)}.items():
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
    bits, _ = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I)
    out[tag + "_points"] = int(masks.sum().item())
    for method, steps in (("pca", 0), ("sweep", 36), ("sweep", 360), ("convex_hull", 0)):
        for _ in range(2):
            rec = ops.fit_all_points(depth, prep, bits, I, torch.float32, method, steps)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            a.record()
            ops.fit_all_points(depth, prep, bits, I, torch.float32, method, steps)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        out[f"{tag}_{method}{steps or ''}_ms"] = round(ts[5], 4)
        out[f"{tag}_{method}{steps or ''}_status_ok"] = bool((rec[..., 41] == 0).all().item())
    print({k: v for k, v in out.items() if k.startswith(tag)}, flush=True)
    del depth, K, masks, ground, bits, prep
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "fitall_methods.json"), "w"), indent=1)
