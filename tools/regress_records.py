"""Bit-level regression vectors of the fit kernels (run on the B200 box).

    python tools/regress_records.py --write PATH.npz           # mint from the build in the tree
    python tools/regress_records.py --write-all PATH.npz       # mint only the all-pixels records (keys */all*)
    python tools/regress_records.py --check PATH.npz [MORE.npz] # compare the build in the tree, bitwise; keys of later files win

The records of `la3d_fit_boxes` / `la3d_fit_boxes_all` on seeded synthetic inputs (small cases: the
float64 records themselves; BASELINE.json's configs at their per-GPU size: a SHA-256 of the float32
records) so that a refactor of the kernels (one shared record tail, new candidate evaluation, ...)
can be asserted bit-identical with the build the parity tests passed on.  The inputs are made on
the device by `synth.make_inputs`; their SHA-256 is stored too, so a box whose torch generator
differs is detected (the check is then skipped for that case, not failed).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SMALL = [  # name, B, I, H, W, method, yaw_steps, ground
    ("s_pca_g", 6, 8, 480, 640, "pca", 0, True),
    ("s_pca_n", 6, 8, 480, 640, "pca", 0, False),
    ("s_hull_g", 6, 8, 480, 640, "convex_hull", 0, True),
    ("s_hull_n", 6, 8, 480, 640, "convex_hull", 0, False),
    ("s_sw36_g", 6, 8, 480, 640, "sweep", 36, True),
    ("s_sw36_n", 6, 8, 480, 640, "sweep", 36, False),
    ("s_sw360_g", 6, 8, 480, 640, "sweep", 360, True),
    ("s_odd_sw7", 3, 5, 97, 131, "sweep", 7, True),
    ("s_odd_hull", 3, 5, 97, 131, "convex_hull", 0, True),
]
FULL = [  # name, B, I, H, W, method, yaw_steps
    ("cfg2_sweep36", 256, 8, 480, 640, "sweep", 36),
    ("cfg2_pca", 256, 8, 480, 640, "pca", 0),
    ("cfg3_pca", 256, 10, 480, 640, "pca", 0),
    ("cfg4_pca", 16, 20, 1536, 1536, "pca", 0),
    ("cfg5_sweep360", 128, 32, 480, 640, "sweep", 360),
    ("cfg2_hull", 64, 8, 480, 640, "convex_hull", 0),
]


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        if t is not None:
            h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def compute(only_all=False):
    from labelany3d_b200 import ops, synth
    out = {}
    for name, B, I, H, W, method, steps, with_g in SMALL:
        d, K, m, g = synth.make_inputs(B, H, W, I, seed=77, device="cuda", area=(0.02, 0.12) if H > 200 else (0.05, 0.3))
        g = g if with_g else None
        out[name + "/in"] = np.array(sha(d, K, m, g))
        if not only_all:
            out[name + "/rec"] = ops.fit_boxes(d, K, m, g, method, steps, seed=1234, out_dtype=torch.float64).cpu().numpy()
            out[name + "/rec32"] = ops.fit_boxes(d, K, m, g, method, steps, seed=1234, out_dtype=torch.float32).cpu().numpy()
        if H > 200:
            out[name + "/all"] = ops.fit_boxes_all(d, K, m, g, out_dtype=torch.float64, method=method, yaw_steps=steps).cpu().numpy()
    for name, B, I, H, W, method, steps in ([] if only_all else FULL):
        d, K, m, g = synth.make_inputs(B, H, W, I, seed=1234 + 2, device="cuda")
        out[name + "/in"] = np.array(sha(d, K, m, g))
        rec = ops.fit_boxes(d, K, m, g, method, steps, seed=1234, out_dtype=torch.float32)
        out[name + "/sha"] = np.array(sha(rec))
        del d, K, m, g
        torch.cuda.empty_cache()
    return out


def main():
    mode, path = sys.argv[1], sys.argv[2]
    got = compute(only_all=(mode == "--write-all"))
    if mode in ("--write", "--write-all"):
        if mode == "--write-all":
            got = {k: v for k, v in got.items() if k.endswith("/all") or k.endswith("/in")}
        np.savez_compressed(path, **got)
        print("wrote", path, os.path.getsize(path), "bytes")
        return 0
    bad = 0
    want = {}
    for pth in sys.argv[2:]:
        with np.load(pth) as z:
            want.update({k: z[k] for k in z.files})
    for key in want:
        if key.endswith("/in"):
            continue
        case = key.split("/")[0]
        if str(want[case + "/in"]) != str(got[case + "/in"]):
            print(f"SKIP {key}: inputs differ on this box (generator mismatch)")
            continue
        a, b = want[key], got[key]
        same = (a.tobytes() == b.tobytes()) if a.dtype.kind == "f" else (str(a) == str(b))
        print(("ok   " if same else "DIFF ") + key)
        if not same and a.dtype.kind == "f":
            with np.errstate(invalid="ignore"):
                diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
            print("      max |diff| =", np.nanmax(diff), "at", np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))[:4].tolist())
        bad += 0 if same else 1
    print("regression:", "IDENTICAL" if bad == 0 else f"{bad} arrays differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
