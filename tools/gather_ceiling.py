"""Ceiling of the in-place depth gather: scattered 4-byte reads from pinned host memory (over PCIe) and from HBM.

    python tools/gather_ceiling.py [OUT.json]

The run-length end-to-end leg leaves the depth maps in pinned host memory and lets the fit kernel read its 500
samples per box in place (2 KB of useful bytes per box instead of a 1.2 MB depth map over the bus).  This tool
measures what that access pattern can reach on the box: reads per second against the loads in flight per thread
and the number of CTAs, for the whole buffer and for mask-sized windows, next to the plain pinned-memory copy rate.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from labelany3d_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    n = 256 * 480 * 640                                   # configs[1]'s depth maps: 315 MB
    host = torch.rand(n, dtype=torch.float32).pin_memory()
    hbm = host.to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    res = {"elements": n, "pinned_copy_GBps": None, "cases": []}

    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tmp = torch.empty_like(hbm)
    tmp.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        tmp.copy_(host, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    res["pinned_copy_GBps"] = round(5 * n * 4 / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
    del tmp

    def run(src, where, window, inflight, ctas, rounds, share=1, spread=0):
        out = torch.empty(ctas * 256, dtype=torch.float32, device=dev)
        args = (src.data_ptr(), n, window, inflight, rounds, ctas, share, spread, out.data_ptr(), stream)
        assert lib.la3d_debug_scatter_read(*args) == 0, _lib.last_error()
        torch.cuda.synchronize()
        a.record()
        assert lib.la3d_debug_scatter_read(*args) == 0
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        reads = ctas * 256 * inflight * rounds
        case = {"memory": where, "window": window, "inflight_per_thread": inflight, "ctas": ctas, "reads": reads,
                "lanes_per_line": share, "elements_apart": spread,
                "ms": round(ms, 3), "Mreads_per_s": round(reads / ms / 1e3, 1)}
        res["cases"].append(case)
        print(case, flush=True)

    for inflight in (1, 4, 16):                            # host memory, anywhere in the buffer
        for ctas in (148, 148 * 8):
            run(host, "pinned host", n, inflight, ctas, max(1, 4_000_000 // (ctas * 256 * inflight)))
    for window in (20_000 * 8, 640 * 200):                 # mask-sized windows (a 200-row band of one image)
        run(host, "pinned host", window, 8, 148 * 8, 2)
    for share, spread in ((2, 1), (2, 8), (2, 16), (4, 8), (32, 1)):   # neighbouring lanes in one 128-byte line
        run(host, "pinned host", n, 8, 148 * 8, 2, share, spread)
    for inflight in (2, 8):                                # the same pattern against HBM
        run(hbm, "hbm", n, inflight, 148 * 8, 64 // inflight * 8)
    best = max(c["Mreads_per_s"] for c in res["cases"] if c["memory"] == "pinned host")
    res["pinned_host_best_Mreads_per_s"] = best
    res["boxes_per_s_at_500_reads"] = round(best * 1e6 / 500)
    print(json.dumps({k: v for k, v in res.items() if k != "cases"}))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
