"""Step time of la3d_fit_boxes for every BASELINE.json config at its per-GPU size (one GPU, device-resident
inputs, CUDA events, float32 records), for both the config's own method and the reference default (pca)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

rows = []
for cfg in (1, 2, 3, 4, 5):
    c = synth.CONFIGS[cfg]
    B, I, H, W = max(c["B"] // c["gpus"], 1), c["I"], c["H"], c["W"]
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + cfg, device="cuda", chunk=4 if cfg == 4 else 16)
    fit = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
    methods = [(c["method"], c["yaw_steps"])]
    if c["method"] != "pca":
        methods.append(("pca", 0))
    for method, steps in methods:
        fn = lambda: fit(depth, K, masks, ground, method, steps, seed=1234)  # noqa: E731
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 50 if cfg != 4 else 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        ok = int((fit.records[..., 41] == 0).sum().item())
        rows.append({"config": cfg, "images_per_gpu": B, "instances": I, "H": H, "W": W, "method": method, "yaw_steps": steps,
                     "ms_per_step": round(ms, 4), "boxes_per_s": round(B * I / ms * 1e3), "mask_GBs": round(B * I * H * W / ms / 1e6, 1),
                     "boxes_ok": ok, "boxes": B * I})
    del depth, masks, fit
    torch.cuda.empty_cache()
print(json.dumps(rows, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/config_table.json", "w"), indent=1)
