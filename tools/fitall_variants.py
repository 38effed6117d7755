"""Time the all-pixels fit on a config-2-shaped batch for both LA3D_FITALL_VARIANT values (CUDA events) and
check that they describe the same boxes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

c = synth.CONFIGS[2]
B, I, H, W = c["B"], c["I"], c["H"], c["W"]
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
out, recs = {}, {}
for v in ("0", "1"):
    os.environ["LA3D_FITALL_VARIANT"] = v
    for _ in range(2):
        recs[v] = ops.fit_boxes_all(depth, K, masks, ground)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.fit_boxes_all(depth, K, masks, ground)
    b.record()
    torch.cuda.synchronize()
    out[f"step_v{v}_ms"] = a.elapsed_time(b) / 10
out["max_abs_diff_between_variants"] = float((recs["0"] - recs["1"]).abs().nan_to_num().max())
out["counts_equal"] = bool(torch.equal(recs["0"][..., 40:42], recs["1"][..., 40:42]))
print(json.dumps(out))
