"""Per-kernel timings on one GPU (CUDA events, inputs larger than L2).  Development aid, not bench.py."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402


def timeit(fn, iters=20, warm=3):
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
    return ts[len(ts) // 2], ts[0]


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    c = dict(synth.CONFIGS[cfg])
    if len(sys.argv) > 2:
        c["B"] = int(sys.argv[2])
    B, I, H, W = c["B"], c["I"], c["H"], c["W"]
    t0 = time.time()
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=1234 + cfg, device="cuda")
    torch.cuda.synchronize()
    out = {"cfg": cfg, "B": B, "I": I, "H": H, "W": W, "gen_s": round(time.time() - t0, 2)}
    px = B * H * W
    lib = _lib.load()

    o32 = torch.empty((B, H, W, 3), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    med, best = timeit(lambda: lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, o32.data_ptr(), 0, st))
    out["lift_f32_ms"] = med; out["lift_f32_GBs"] = px * 16 / med / 1e6; out["lift_f32_best_GBs"] = px * 16 / best / 1e6
    del o32
    if px * 24 < 40e9:
        o64 = torch.empty((B, H, W, 3), dtype=torch.float64, device="cuda")
        med, best = timeit(lambda: lib.la3d_depth_lift(depth.data_ptr(), K.data_ptr(), 9, 0, None, None, B, H, W, o64.data_ptr(), 1, st))
        out["lift_f64_ms"] = med; out["lift_f64_GBs"] = px * 28 / med / 1e6
        del o64

    m8 = masks.view(torch.uint8)
    chunks, words = ops.scan_layout(H, W)
    bits = torch.empty((B * I, words), dtype=torch.int32, device="cuda")
    cc = torch.empty((B * I, chunks), dtype=torch.int32, device="cuda")
    for is01 in (1, 0):
        med, best = timeit(lambda: lib.la3d_mask_scan(m8.data_ptr(), B * I, H, W, is01, bits.data_ptr(), cc.data_ptr(), st))
        out[f"scan{is01}_ms"] = med; out[f"scan{is01}_GBs"] = px * I / med / 1e6; out[f"scan{is01}_best_GBs"] = px * I / best / 1e6
    counts = torch.empty((B, I), dtype=torch.int32, device="cuda")
    ranks = torch.empty((B, I, 500), dtype=torch.int32, device="cuda")
    prep = torch.empty(lib.la3d_prep_bytes(B, I), dtype=torch.uint8, device="cuda")
    med, best = timeit(lambda: lib.la3d_fit_prepare(K.data_ptr(), ground.data_ptr(), B, I, 1234, 0, prep.data_ptr(), prep.numel(), st))
    out["prepare_ms"] = med
    med, best = timeit(lambda: lib.la3d_sample_ranks(cc.data_ptr(), prep.data_ptr(), B, I, H, W, counts.data_ptr(), ranks.data_ptr(), st))
    out["sample_ms"] = med
    rec = torch.empty((B, I, 64), dtype=torch.float64, device="cuda")
    for name, mid, steps in (("pca", 0, 0), ("hull", 1, 0), ("sweep36", 2, 36), ("sweep360", 2, 360)):
        med, best = timeit(lambda: lib.la3d_fit_scanned(depth.data_ptr(), prep.data_ptr(), bits.data_ptr(), cc.data_ptr(),
                                                        ranks.data_ptr(), B, I, H, W, mid, steps, rec.data_ptr(), 1, st))
        out[f"fit_{name}_ms"] = med
    # "next" rows: mask statistics, overlap, masked-median depth scale
    stats = torch.empty((B * I, 8), dtype=torch.int32, device="cuda")
    med, best = timeit(lambda: ops.mask_stats(bits, H, W, 10))
    out["mask_stats_ms"] = med; out["mask_stats_GBs"] = bits.numel() * 4 / med / 1e6
    fbits = bits.view(B, I, words)[:, 0].contiguous()
    med, best = timeit(lambda: ops.mask_overlap(bits, fbits, H, W, group=I))
    out["mask_overlap_ms"] = med; out["mask_overlap_GBs"] = bits.numel() * 4 * (1 + 1.0 / I) / med / 1e6
    if px * I * 4 < 20e9:
        render = (depth[:, None] / 2.5).expand(B, I, H, W).contiguous()
        rb, _ = ops.mask_scan(torch.roll(masks, shifts=(3, -5), dims=(2, 3)))     # a rendered object overlaps most of its mask
        med, best = timeit(lambda: ops.depth_scale_median(depth, render, bits, rb, H, W), iters=5, warm=1)
        n_ov, _ = ops.depth_scale_median(depth, render, bits, rb, H, W)
        out["ratio_median_ms"] = med; out["ratio_median_overlap_px"] = int(n_ov.sum().item())
        del render
    fit = ops.BoxFitter(B, I, H, W)
    for name, steps in (("pca", 0), ("sweep", c["yaw_steps"] or 36)):
        med, best = timeit(lambda: fit(depth, K, masks, ground, name, steps, seed=1234))
        out[f"e2e_dev_{name}_ms"] = med; out[f"e2e_dev_{name}_boxes_per_s"] = B * I / med * 1e3
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/microbench_cfg{cfg}.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
