"""Time the pieces of the run-length path on a config-2-shaped batch with CUDA events (20 launches each):
stand-alone preparation, plain decode and the whole la3d_fit_boxes_rle step for each LA3D_RLE_VARIANT,
and the sampler / fit kernels on the decoded planes."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, coco_rle, ops, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
B, I, H, W = c["B"], c["I"], c["H"], c["W"]
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + cfg, device="cuda")
host = masks.cpu().numpy().reshape(B * I, H, W)
counts, offsets, max_runs = coco_rle.pack_runs([coco_rle.runs_from_mask(m) for m in host])
d_counts = torch.as_tensor(counts.view(np.int32), device="cuda")
d_off = torch.as_tensor(offsets, device="cuda")
lib = _lib.load()
fitter = ops.RleBoxFitter(B, I, H, W, d_counts.numel(), max_runs, out_dtype=torch.float32)
N = 20


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(N):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / N * 1e3          # microseconds per launch, back to back


out = {"config": cfg, "runs": int(d_counts.numel()), "max_runs": max_runs}
out["prepare_us"] = timed(lambda: ops.fit_prepare(K, ground, B, I, 1234, 0))
for v in ("0", "1"):
    os.environ["LA3D_RLE_VARIANT"] = v
    out[f"decode_plain_v{v}_us"] = timed(lambda: ops.rle_decode(d_counts, d_off, H, W, max_runs))
    out[f"step_rle_v{v}_us"] = timed(lambda: fitter(depth, K, d_counts, d_off, ground, "sweep", c["yaw_steps"] or 36, seed=1234))
    out[f"step_rle_pca_v{v}_us"] = timed(lambda: fitter(depth, K, d_counts, d_off, ground, "pca", 0, seed=1234))
os.environ["LA3D_RLE_VARIANT"] = "0"
bits, cc, _ = ops.rle_decode(d_counts, d_off, H, W, max_runs)
prep = ops.fit_prepare(K, ground, B, I, 1234, 0)
out["sample_us"] = timed(lambda: ops.sample_ranks(cc, B, I, H, W, prep=prep))
_, ranks = ops.sample_ranks(cc, B, I, H, W, prep=prep)
out["fit_sweep_us"] = timed(lambda: ops.fit_scanned(depth, prep, bits, cc, ranks, "sweep", c["yaw_steps"] or 36, out_dtype=torch.float32))
out["fit_pca_us"] = timed(lambda: ops.fit_scanned(depth, prep, bits, cc, ranks, "pca", 0, out_dtype=torch.float32))
out["scan_us"] = timed(lambda: ops.mask_scan(masks))
out["note"] = "ops.* wrappers allocate their outputs per call (torch caching allocator); back-to-back launches"
print(json.dumps(out))
