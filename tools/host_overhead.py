"""Host-side cost of one step: tiny batch (the GPU work is negligible), plain BoxFitter call against the sharded form's
extra work (status check, la3d_sink construction, input slicing)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

B, I, H, W = 1, 8, 64, 64
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda", area=(0.05, 0.3))
fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
bufs = [torch.empty((8 * B, I, 64), dtype=torch.float32, device="cuda") for _ in range(8)]
flags = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(8)]
counter = torch.zeros(1, dtype=torch.int32, device="cuda")
status = torch.zeros(1, dtype=torch.int32).pin_memory()
for f in flags:
    f.fill_(1 << 30)            # every "peer" is far ahead: no waiting
epoch = [0]


def plain():
    fitter(depth, K, masks, ground, "sweep", 36, seed=1)


def sharded_like():
    if int(status[0]) != 0:
        raise RuntimeError
    epoch[0] += 1
    sink = _lib.make_sink([b.data_ptr() for b in bufs], False, [f.data_ptr() for f in flags], counter.data_ptr(), status.data_ptr(), epoch[0], 0)
    fitter(depth[:B], K[:B], masks[:B], ground[:B], "sweep", 36, seed=1, sink=sink)


for name, fn in (("plain BoxFitter call", plain), ("sharded-like call (8 destinations)", sharded_like)):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    torch.cuda.synchronize()
    print(f"{name:36s} {(time.perf_counter() - t0) / 2000 * 1e6:6.1f} us per step (host-bound)", flush=True)
