"""Sampler and whole-step time against the words a sampler CTA stages at a time (la3d_set_sample_seg_blocks), one
process: 2048 and 256 images of configs[1]'s shape; records compared bitwise with the default's.  argv: OUT.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
res = {}
for B in (2048, 256):
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1236, device="cuda")
    fit = ops.BoxFitter(B, I, H, W, out_dtype=torch.float32)
    want = None
    for seg in (0, 12, 10, 8, 6, 5, 4, 3):
        lib.la3d_set_sample_seg_blocks(seg)
        for _ in range(3):
            rec = fit(depth, K, masks, ground, "sweep", 36, seed=1234)
        if want is None:
            want = rec.clone()
        same = bool(torch.equal(rec.view(torch.int32), want.view(torch.int32)))
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(20)]
        torch.cuda.synchronize()
        for e in evs:
            fit(depth, K, masks, ground, "sweep", 36, seed=1234, events=e)
        torch.cuda.synchronize()
        med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fit(depth, K, masks, ground, "sweep", 36, seed=1234)
        b.record()
        torch.cuda.synchronize()
        res[f"B{B}_seg{seg}"] = {"sample_us": round(med([e[1].elapsed_time(e[2]) for e in evs]) * 1e3, 1),
                                  "step_us": round(a.elapsed_time(b) / 20 * 1e3, 1), "identical": same}
        print(f"B{B}_seg{seg}", res[f"B{B}_seg{seg}"], flush=True)
    lib.la3d_set_sample_seg_blocks(0)
    del depth, K, masks, ground, fit
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
