"""The run-length end-to-end leg of bench.py alone (one GPU): boxes/s with the runs, intrinsics and ground normals
copied in from pinned host memory every step, the depth read in place, records back to pinned host memory.

    LA3D_FIT_SORT_DEPTH=0 python tools/rle_e2e.py     # the unsorted in-place gather, for comparison
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    images = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    args = types.SimpleNamespace(steps=10, warmup=3)
    cx = bench.Ctx(args)
    import torch
    from labelany3d_b200 import synth
    torch.cuda.set_device(cx.dev)
    w = bench.SHAPE
    depth, K, masks, ground = synth.make_inputs(images, w["H"], w["W"], w["I"], seed=bench.SEED, device=cx.dev)
    leg = bench.rle_leg(cx, depth, K, masks, ground, images, 0, images, "none", 10, 3)
    print(json.dumps({"images": images, "sort": os.environ.get("LA3D_FIT_SORT_DEPTH", "1"),
                      "identical": leg["records_identical_to_byte_mask_path"],
                      "device_boxes_per_s": round(leg["value"]), "e2e_boxes_per_s": round(leg["e2e"]["value"]),
                      "e2e_ms_per_step": round(leg["e2e"]["ms_per_step"], 3)}))


if __name__ == "__main__":
    main()
