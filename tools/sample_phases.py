"""Timeline of the sampler's CTAs (la3d_debug_sample_clocks: globaltimer stamps): when each image's CTA starts, how long
the counting, the staging wait and the instance walk take, against the kernel's CUDA-event duration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

lib = _lib.load()
for B in (256, 2048):
    I, H, W = 8, 480, 640
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
    bits, cc = ops.mask_scan(masks)
    prep = ops.fit_prepare(K, ground, B, I, seed=1)
    counts, ranks = ops.sample_ranks(cc, B, I, H, W, prep=prep)
    st = torch.cuda.current_stream().cuda_stream
    clocks = torch.zeros((B, 4), dtype=torch.int64, device="cuda")
    call = lambda: lib.la3d_sample_ranks(cc.data_ptr(), prep.data_ptr(), B, I, H, W, counts.data_ptr(), ranks.data_ptr(), st)  # noqa: E731
    for _ in range(3):
        call()
    lib.la3d_debug_sample_clocks(clocks.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); call(); b.record()
    torch.cuda.synchronize()
    lib.la3d_debug_sample_clocks(None)
    c = clocks.double()
    t0 = c[:, 0].min()
    print(f"B={B}: kernel {a.elapsed_time(b) * 1e3:.1f} us (events); CTA starts spread {float(c[:, 0].max() - t0) / 1e3:.1f} us; "
          f"per CTA mean: counts {float((c[:, 1] - c[:, 0]).mean()) / 1e3:.2f} us, staging wait {float((c[:, 2] - c[:, 1]).mean()) / 1e3:.2f} us, "
          f"walk {float((c[:, 3] - c[:, 2]).mean()) / 1e3:.2f} us (max {float((c[:, 3] - c[:, 2]).max()) / 1e3:.2f}); "
          f"last CTA ends {float(c[:, 3].max() - t0) / 1e3:.1f} us after the first starts", flush=True)
