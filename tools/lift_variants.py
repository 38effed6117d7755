"""Times the depth-lift kernel variants (LA3D_LIFT_VARIANT is read once per process: run per variant)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
B, H, W = c["B"], c["H"], c["W"]
depth, K, _, _ = synth.make_inputs(B, H, W, 1, seed=1, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
px = B * H * W
res = {"variant": os.environ.get("LA3D_LIFT_VARIANT", "0"), "cfg": cfg}
for f64, dt, bpp in ((0, torch.float32, 16), (1, torch.float64, 28)):
    out = torch.empty((B, H, W, 3), dtype=dt, device="cuda")
    fn = lambda: lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, out.data_ptr(), f64, st)  # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    res["f64" if f64 else "f32"] = {"ms": round(ms, 4), "GBs": round(px * bpp / ms / 1e6, 1)}
    res["sum%d" % f64] = float(out[0, :4, :4].double().sum())
    del out
print(json.dumps(res))
