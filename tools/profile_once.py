"""Launch each kernel of the path twice on a config-2-shaped batch (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
if len(sys.argv) > 2:
    c["B"] = int(sys.argv[2])
which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["lift", "scan", "sample", "pca", "sweep", "hull", "step", "next"]
B, I, H, W = c["B"], c["I"], c["H"], c["W"]
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + cfg, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
m8 = masks.view(torch.uint8)
chunks, words = ops.scan_layout(H, W)
bits = torch.empty((B * I, words), dtype=torch.int32, device="cuda")
cc = torch.empty((B * I, chunks), dtype=torch.int32, device="cuda")
counts = torch.empty((B, I), dtype=torch.int32, device="cuda")
ranks = torch.empty((B, I, 500), dtype=torch.int32, device="cuda")
rec = torch.empty((B, I, 64), dtype=torch.float64, device="cuda")
o32 = torch.empty((B, H, W, 3), dtype=torch.float32, device="cuda")
prep = torch.empty(lib.la3d_prep_bytes(B, I), dtype=torch.uint8, device="cuda")
fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
torch.cuda.synchronize()
for _ in range(2):
    if "lift" in which:
        lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, o32.data_ptr(), 0, st)
    lib.la3d_fit_prepare(K.data_ptr(), ground.data_ptr(), B, I, 1234, 0, prep.data_ptr(), prep.numel(), st)
    lib.la3d_mask_scan(m8.data_ptr(), B * I, H, W, 1, bits.data_ptr(), cc.data_ptr(), st)
    lib.la3d_sample_ranks(cc.data_ptr(), prep.data_ptr(), B, I, H, W, counts.data_ptr(), ranks.data_ptr(), st)
    for name, mid, steps in (("pca", 0, 0), ("sweep", 2, c["yaw_steps"] or 36), ("hull", 1, 0)):
        if name in which:
            lib.la3d_fit_scanned(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(),
                                 ranks.data_ptr(), B, I, H, W, mid, steps, rec.data_ptr(), 1, st)
    if "step" in which:          # the one-call pipeline as bench.py runs it: scan (+ prep CTAs), sampler, fit
        fitter(depth, K, masks, ground, "sweep", c["yaw_steps"] or 36, seed=1234)
    if "next" in which:
        lib.la3d_mask_scan_thin(m8.data_ptr(), B * I, H, W, 1, bits.data_ptr(), cc.data_ptr(), 3, 2, st)
        ops.mask_stats(bits, H, W, 10)
        ops.mask_overlap(bits, bits.view(B, I, words)[:, 0].contiguous(), H, W, group=I)
        if B * I * H * W * 4 < 20e9:
            render = (depth[:, None] / 2.5).expand(B, I, H, W).contiguous()
            ops.depth_scale_median(depth, render, bits, torch.roll(bits, 1, 0).contiguous(), H, W)
            del render
torch.cuda.synchronize()
print("done")
