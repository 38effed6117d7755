"""Times la3d_mask_scan alone (LA3D_SCAN_VARIANT / LA3D_SCAN_CTAS are read once per process)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from labelany3d_b200 import _lib, ops, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synth.CONFIGS[cfg])
B, I, H, W = c["B"] // c["gpus"], c["I"], c["H"], c["W"]
masks = (torch.rand((B, I, H, W), device="cuda") < 0.1)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
chunks, words = ops.scan_layout(H, W)
bits = torch.empty((B * I, words), dtype=torch.int32, device="cuda")
cc = torch.empty((B * I, chunks), dtype=torch.int32, device="cuda")
m8 = masks.view(torch.uint8)
fn = lambda: lib.la3d_mask_scan(m8.data_ptr(), B * I, H, W, 1, bits.data_ptr(), cc.data_ptr(), st)  # noqa: E731
for _ in range(5):
    fn()
torch.cuda.synchronize()
n = 30
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    fn()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
ok = int(cc.view(torch.uint8).sum().item()) == int(masks.sum().item())
print(json.dumps({"variant": os.environ.get("LA3D_SCAN_VARIANT", "0"), "ctas": os.environ.get("LA3D_SCAN_CTAS", "1"), "cfg": cfg,
                  "ms": round(ms, 4), "GBs": round(B * I * H * W / ms / 1e6, 1), "counts_ok": ok}))
