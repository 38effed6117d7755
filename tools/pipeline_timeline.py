"""Two batches in flight on two streams (scan of batch k+1 under the sampler + fit of batch k): when do the
kernels run and how long do they take when they share the SMs?  The experiment behind
profiles/r1_pipeline_timeline.txt (argv: CTAs per SM and stages of the thin scan; 0 0 = the tile scan)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

B, I, H, W = 256, 8, 480, 640
ctas, stages = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 3)
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
lib = _lib.load()
m8 = masks.view(torch.uint8)
slots = [ops.BoxFitter(B, I, H, W, out_dtype=torch.float32) for _ in range(2)]
prio = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # -1: the sampler / fit stream gets the higher priority
s_scan, s_fit = torch.cuda.Stream(), torch.cuda.Stream(priority=prio)
N = 8
E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
ev = [{k: E() for k in ("scan0", "scan1", "prep1", "samp0", "samp1", "fit1")} for _ in range(N)]
done = [None, None]
t0 = E()
for rep in range(2):          # first repetition warms up
    torch.cuda.synchronize()
    t0.record()
    s_scan.wait_event(t0)
    s_fit.wait_event(t0)
    for k in range(N):
        f = slots[k & 1]
        bits, cc, counts, ranks, prep, prep_bytes = f._carve()
        e = ev[k]
        if done[k & 1] is not None:
            s_scan.wait_event(done[k & 1])
        e["scan0"].record(s_scan)
        if ctas > 0:
            lib.la3d_mask_scan_thin(m8.data_ptr(), B * I, H, W, 1, bits, cc, ctas, stages, s_scan.cuda_stream)
        else:
            lib.la3d_mask_scan(m8.data_ptr(), B * I, H, W, 1, bits, cc, s_scan.cuda_stream)
        e["scan1"].record(s_scan)
        lib.la3d_fit_prepare(K.data_ptr(), ground.data_ptr(), B, I, 1234, 0, prep, prep_bytes, s_fit.cuda_stream)
        e["prep1"].record(s_fit)
        s_fit.wait_event(e["scan1"])
        e["samp0"].record(s_fit)
        lib.la3d_sample_ranks(cc, prep, B, I, H, W, counts, ranks, s_fit.cuda_stream)
        e["samp1"].record(s_fit)
        lib.la3d_fit_scanned(depth.data_ptr(), prep, bits, cc, ranks, B, I, H, W, 2, 36, f.records.data_ptr(), 0, s_fit.cuda_stream)
        e["fit1"].record(s_fit)
        done[k & 1] = e["fit1"]
    torch.cuda.synchronize()
print(f"ctas={ctas} stages={stages} tail-stream priority={prio}: microseconds since start")
for k in range(N):
    e = ev[k]
    us = {n: t0.elapsed_time(x) * 1e3 for n, x in e.items()}
    print(f"batch {k}: scan {us['scan0']:7.1f} -> {us['scan1']:7.1f} ({us['scan1'] - us['scan0']:6.1f}) | prep done {us['prep1']:7.1f} | "
          f"sample {us['samp0']:7.1f} -> {us['samp1']:7.1f} ({us['samp1'] - us['samp0']:5.1f}) | fit -> {us['fit1']:7.1f} ({us['fit1'] - us['samp1']:6.1f})")
