"""Launch the depth-lift kernels a few times on a config-2 batch (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
if len(sys.argv) > 2:
    c["B"] = int(sys.argv[2])
depth, K, masks, ground = synth.make_inputs(c["B"], c["H"], c["W"], 1, seed=1, device="cuda")
for _ in range(3):
    ops.depth_lift(depth, K, out_dtype=torch.float32)
    ops.depth_lift(depth, K, out_dtype=torch.float64)
torch.cuda.synchronize()
print("done")
