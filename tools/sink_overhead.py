"""In-kernel cost of the record sink's synchronisation on ONE GPU: the step with (a) the plain local record buffer,
(b) a sink of two local destinations without flags, (c) the same with the flag protocol (both flag rows local, the
"peer" epoch published by la3d_peer_signal each step)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
B, I, H, W = 256, 8, 480, 640
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
fitter = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
bufs = [torch.empty((2 * B, I, 64), dtype=torch.float32, device="cuda") for _ in range(2)]
flags = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(2)]
counter = torch.zeros(1, dtype=torch.int32, device="cuda")
status = torch.zeros(1, dtype=torch.int32).pin_memory()
arr = (ctypes.c_void_p * 2)(*[f.data_ptr() for f in flags])
epoch = [0]


def plain():
    fitter(depth, K, masks, ground, "sweep", 36, seed=1)


def two_dest():
    fitter(depth, K, masks, ground, "sweep", 36, seed=1, sink=_lib.make_sink([b.data_ptr() for b in bufs], False))


def synced():
    epoch[0] += 1
    e = epoch[0]
    sink = _lib.make_sink([b.data_ptr() for b in bufs], False, [f.data_ptr() for f in flags], counter.data_ptr(), status.data_ptr(), e, 0)
    fitter(depth, K, masks, ground, "sweep", 36, seed=1, sink=sink)
    flags[0][1:2].fill_(e)                                                          # the "peer" keeps up (one tiny fill)


for name, fn in (("plain", plain), ("two destinations", two_dest), ("two destinations + flags", synced), ("plain", plain)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {a.elapsed_time(b) / 40 * 1e3:7.1f} us per step", flush=True)
