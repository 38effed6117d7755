"""Step time of the one-call pipeline: eager launches vs CUDA-graph replay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import ops, synth  # noqa: E402

B, I, H, W = 256, 8, 480, 640
method, steps = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("sweep", 36)
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
fit = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)


def timeit(fn, n=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


eager = timeit(lambda: fit(depth, K, masks, ground, method, steps, seed=1234))
ref = fit(depth, K, masks, ground, method, steps, seed=1234).clone()
replay, rec = fit.capture(depth, K, masks, ground, method, steps, seed=1234)
graph = timeit(replay)
torch.cuda.synchronize()
same = torch.equal(torch.nan_to_num(rec), torch.nan_to_num(ref))
print(f"{method}{steps}: eager {eager:.1f} us/step, graph {graph:.1f} us/step, identical={same}")
