"""la3d_depth_lift alone (float32 and float64 points) at BASELINE configs[1] and configs[3]: CUDA events over 20
back-to-back calls, bytes = 16 (28) per pixel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, synth  # noqa: E402

lib = _lib.load()
out = {}
for name, (B, H, W) in {"cfg2_256x640x480": (256, 480, 640), "cfg4_128x1536x1536": (128, 1536, 1536)}.items():
    depth, K, _, _ = synth.make_inputs(2, H, W, 1, seed=7, device="cuda")
    depth = depth[:1].expand(B, H, W).contiguous()
    K = K[:1].expand(B, 3, 3).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    for f64, bpp in ((0, 16), (1, 28)):
        pts = torch.empty((B, H, W, 3), dtype=torch.float64 if f64 else torch.float32, device="cuda")
        call = lambda: lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, pts.data_ptr(), f64, st)  # noqa: E731
        for _ in range(3):
            call()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            call()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        out[f"{name}_{'f64' if f64 else 'f32'}"] = {"ms": round(ms, 4), "GBps": round(B * H * W * bpp / ms / 1e6, 1)}
        del pts
    print(name, {k: v for k, v in out.items() if k.startswith(name)}, flush=True)
json.dump(out, open(os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "lift_bench.json"), "w"), indent=1)
