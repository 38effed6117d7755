"""Timeline of a 2-part pipeline (scan on the main stream, sample + fit on a high-priority side stream)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

B, I, H, W = 256, 8, 480, 640
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prio = int(sys.argv[2]) if len(sys.argv) > 2 else -1
depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=3, device="cuda")
lib = _lib.load()
m8 = masks.view(torch.uint8)
chunks, words = ops.scan_layout(H, W)
bits = torch.empty((B * I, words), dtype=torch.int32, device="cuda")
cc = torch.empty((B * I, chunks), dtype=torch.int32, device="cuda")
counts = torch.empty((B, I), dtype=torch.int32, device="cuda")
ranks = torch.empty((B, I, 500), dtype=torch.int32, device="cuda")
rec = torch.empty((B, I, 64), dtype=torch.float32, device="cuda")
main = torch.cuda.Stream()
sides = [torch.cuda.Stream(priority=prio) for _ in range(parts)]
HW = H * W
nb = B // parts


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(record):
    marks = {}
    with torch.cuda.stream(main):
        t0 = ev(); t0.record(); marks["t0"] = t0
        for s in range(parts):
            b0 = s * nb
            pl0 = b0 * I
            lib.la3d_mask_scan(m8.data_ptr() + pl0 * HW, nb * I, H, W, 1, bits.data_ptr() + pl0 * words * 4, cc.data_ptr() + pl0 * chunks * 4, main.cuda_stream)
            e = ev(); e.record(main); marks[f"scan{s}_end"] = e
            sides[s].wait_event(e)
            with torch.cuda.stream(sides[s]):
                a = ev(); a.record(); marks[f"sample{s}_start"] = a
                lib.la3d_sample_ranks(cc.data_ptr() + pl0 * chunks * 4, nb, I, H, W, 1234, b0, counts.data_ptr() + pl0 * 4, ranks.data_ptr() + pl0 * 2000, sides[s].cuda_stream)
                a = ev(); a.record(); marks[f"sample{s}_end"] = a
                lib.la3d_fit_scanned(depth.data_ptr() + b0 * HW * 4, K.data_ptr() + b0 * 72, ground.data_ptr() + pl0 * 24, bits.data_ptr() + pl0 * words * 4,
                                     cc.data_ptr() + pl0 * chunks * 4, counts.data_ptr() + pl0 * 4, ranks.data_ptr() + pl0 * 2000, nb, I, H, W, 2, 36,
                                     rec.data_ptr() + pl0 * 256, 0, sides[s].cuda_stream)
                a = ev(); a.record(); marks[f"fit{s}_end"] = a
        for s in range(parts):
            main.wait_event(marks[f"fit{s}_end"])
        t1 = ev(); t1.record(main); marks["t1"] = t1
    return marks


for _ in range(3):
    run(False)
torch.cuda.synchronize()
m = run(True)
torch.cuda.synchronize()
t0 = m["t0"]
for k, e in m.items():
    print(f"{k:16s} {t0.elapsed_time(e) * 1e3:8.1f} us")
