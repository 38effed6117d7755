"""The tails of a step alone (sampler, fit kernel per method) on outputs of the scan kept resident: CUDA events,
median of 20 launches.  Development aid for the fit kernel (argv: output JSON name under gpurun_out/)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()


def med(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return round(ts[len(ts) // 2] * 1e3, 1)


out = {}
cases = [(256, 8, 480, 640), (2048, 8, 480, 640), (128, 32, 480, 640), (128, 20, 1536, 1536)]
if len(sys.argv) > 2:
    cases = cases[:int(sys.argv[2])]
for B, I, H, W in cases:
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
    bits, cc = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I, seed=1)
    counts, ranks = ops.sample_ranks(cc, B, I, H, W, prep=prep)
    rec = torch.empty((B, I, 64), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    key = f"B{B}xI{I}x{H}"
    out[key + "_sample"] = med(lambda: lib.la3d_sample_ranks(cc.data_ptr(), prep.data_ptr(), B, I, H, W, counts.data_ptr(), ranks.data_ptr(), st))
    for name, mid, steps in (("pca", 0, 0), ("sweep36", 2, 36), ("sweep360", 2, 360), ("hull", 1, 0)):
        out[f"{key}_{name}"] = med(lambda: lib.la3d_fit_scanned(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(),
                                                                ranks.data_ptr(), B, I, H, W, mid, steps, rec.data_ptr(), 0, st))
    print({k: v for k, v in out.items() if k.startswith(key)}, flush=True)
    del depth, K, masks, ground, bits, cc, prep, counts, ranks, rec
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "fit_bench.json"), "w"), indent=1)
