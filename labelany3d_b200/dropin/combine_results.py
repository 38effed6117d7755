"""Drop-in for the reference's ``src/tools/combine_results.py``: per-scene 3D box results -> one
Omni3D-format JSON.  Same function names, signatures, skip rules, printed warnings and output
schema; the arithmetic of ALL scenes runs in two kernel launches (``la3d_box2d_from_corners`` for
``bbox2D_proj`` / ``bbox2D_trunc``, ``la3d_iou_matrix`` for the Hungarian cost matrices); the
assignment itself stays ``scipy.optimize.linear_sum_assignment`` on the host, like the reference.

    python combine_results.py --split val --results_dir ../experimental_results/COCO
"""

from __future__ import annotations

import argparse
import json
import os

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from labelany3d_b200 import ops as _ops

# Omni3D-style category ids the reference writes (``src/tools/combine_results.py:17-100``): an
# interface constant of the output format, kept as data.
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "coco_omni3d_categories.json")) as _f:
    COCO_CATEGORIES = json.load(_f)
CATEGORY_NAME_TO_ID = {cat["name"]: cat["id"] for cat in COCO_CATEGORIES}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("labelany3d_b200 needs a CUDA device: this path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def _dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=_device())


def project_to_2d(point_3d, K):
    """``(K p)[:2] / (K p)[2]`` (``:105-108``)."""
    uv = _ops.project_points(_dev(np.asarray(point_3d, dtype=np.float64).reshape(-1, 3), np.float64), _dev(K, np.float64))
    return uv.cpu().numpy()[0]


def iou2D(box1, box2):
    """IoU of two xyxy boxes (``:111-123``)."""
    m = _ops.iou_matrix(_dev(np.asarray(box1, dtype=np.float64).reshape(1, 4), np.float64),
                        _dev(np.asarray(box2, dtype=np.float64).reshape(1, 4), np.float64))
    return m.cpu().numpy()[0, 0]


# constant fields of an Omni3D annotation as the reference writes them (:254-262), in its key order
_ANNO_HEAD = (("behind_camera", False), ("truncation", 0.0), ("visibility", 1), ("segmentation_pts", -1),
              ("lidar_pts", -1), ("valid3D", True))
_ANNO_COPIED = ("center_cam", "dimensions", "R_cam", "bbox3D_cam")     # taken over from the 3dbbox JSON entry


def _annotation(src, name, cat_id, image_id, anno_id, dataset_id):
    out = dict(_ANNO_HEAD)
    out.update(category_name=name, category_id=cat_id, image_id=image_id, id=anno_id, dataset_id=dataset_id)
    for key in _ANNO_COPIED:
        out[key] = src.get(key)
    out.update(bbox2D_proj=None, bbox2D_trunc=None, depth_error=-1)      # the two boxes are filled in after the GPU pass
    return out


def _image_entry(scene_name, split, K, W, H, image_id, dataset_id):
    return dict(width=int(W), height=int(H), file_path=f"coco/images/{split}2017/{scene_name}.jpg", K=K.tolist(),
                src_90_rotate=0, src_flagged=False, incomplete=False, id=image_id, dataset_id=dataset_id)


def _match(iou):
    rows, cols = linear_sum_assignment(-iou)
    return [(i, j, iou[i, j]) for i, j in zip(rows, cols)]


def hungarian_matching(boxes0, boxes1):
    """``[(index0, index1, IoU)]`` of the IoU-optimal assignment (``:126-144``)."""
    n0, n1 = len(boxes0), len(boxes1)
    if n0 == 0 or n1 == 0:
        return []
    iou = _ops.iou_matrix(_dev(np.asarray(boxes0, dtype=np.float64).reshape(n0, 4), np.float64),
                          _dev(np.asarray(boxes1, dtype=np.float64).reshape(n1, 4), np.float64)).cpu().numpy()
    return _match(iou)


def combine_coco_results(results_dir, split, output_path, bbox_filename="3dbbox.json"):
    """Combine per-scene 3D bbox results into Omni3D format JSON (``:147-311``)."""
    scene_dir = os.path.join(results_dir, split)
    if not os.path.exists(scene_dir):
        raise FileNotFoundError(f"Results directory not found: {scene_dir}")
    scene_ids = sorted([d for d in os.listdir(scene_dir) if os.path.isdir(os.path.join(scene_dir, d))])
    print(f"Found {len(scene_ids)} scenes in {scene_dir}")

    is_val = split == "val"
    dataset_id = 22 if is_val else 23
    image_id, annotation_id = (1000000, 100000000) if is_val else (2000000, 200000000)

    def warn(text):
        print(f"Warning: {text}")

    def load_json(path):
        with open(path, "r") as f:
            return json.load(f)

    # ---- pass 1 (host): read the scenes, apply the reference's skip rules (:176-231), collect the numbers
    images, scenes = [], []          # scenes: one dict per kept image (annos, 2D boxes, image size)
    corners, k_index, Ks, whs = [], [], [], []
    for scene_name in scene_ids:
        here = os.path.join(scene_dir, scene_name)
        files = {key: os.path.join(here, name) for key, name in
                 (("boxes3d", bbox_filename), ("camera", "cam_params.json"), ("boxes2d", "bboxes.json"))}
        if not os.path.exists(files["boxes3d"]):
            warn(f"Missing {bbox_filename} in {scene_name}, skipping")
            continue
        if not os.path.exists(files["camera"]):
            warn(f"Missing cam_params.json in {scene_name}, skipping")
            continue
        camera = load_json(files["camera"])
        K, H, W = np.array(camera["K"]), camera["H"], camera["W"]
        entries = load_json(files["boxes3d"])
        if len(entries) == 0:
            warn(f"Empty bbox in {scene_name}, skipping")
            continue
        tight = load_json(files["boxes2d"]) if os.path.exists(files["boxes2d"]) else None
        if tight is None:
            warn(f"Missing bboxes.json in {scene_name}, using projected bbox as bbox2D_tight")
        images.append(_image_entry(scene_name, split, K, W, H, image_id, dataset_id))
        local = []
        for entry in entries:
            name = entry.get("category_name", "").replace("_", " ")
            cat_id = CATEGORY_NAME_TO_ID.get(name, -1)
            if cat_id == -1:
                warn(f"Unknown category '{name}' in {scene_name}, skipping")
                continue
            corners.append(np.array(entry["bbox3D_cam"], dtype=np.float64).reshape(8, 3))
            k_index.append(len(Ks))
            local.append(_annotation(entry, name, cat_id, image_id, annotation_id, dataset_id))
            annotation_id += 1
        Ks.append(K.astype(np.float64).reshape(3, 3))
        whs.append([float(W), float(H)])
        scenes.append({"annos": local, "bbox2d": tight, "W": W, "H": H, "match": False})
        image_id += 1

    # ---- pass 2 (GPU): every box of every scene in one launch, every cost matrix in another
    if corners:
        proj, trunc = _ops.box2d_from_corners(_dev(np.stack(corners), np.float64), _dev(np.stack(Ks), np.float64),
                                              _dev(np.array(whs), np.float64), _dev(np.array(k_index), np.int32))
        proj, trunc = proj.cpu().numpy(), trunc.cpu().numpy()
    cursor = 0
    match_scenes, b0, b1, off0, off1 = [], [], [], [0], [0]
    for sc in scenes:
        n = len(sc["annos"])
        for j, anno in enumerate(sc["annos"]):
            p = [float(x) for x in proj[cursor + j]]
            anno["bbox2D_proj"] = p
            # the clamp on the host over the GPU's numbers, so that a clamped entry is the reference's
            # int 0 / W / H (``max(0, x)``, ``min(W, x)`` of :247-252) in the JSON; same values as `trunc`
            anno["bbox2D_trunc"] = [max(0, p[0]), max(0, p[1]), min(sc["W"], p[2]), min(sc["H"], p[3])]
        if sc["bbox2d"] is not None and n > 0 and len(sc["bbox2d"]) > 0:
            sc["match"] = True
            match_scenes.append(sc)
            b0.append(trunc[cursor:cursor + n])
            b1.append(np.array(sc["bbox2d"], dtype=np.float64).reshape(-1, 4))
            off0.append(off0[-1] + n)
            off1.append(off1[-1] + len(sc["bbox2d"]))
        cursor += n
    if match_scenes:
        flat, out_off = _ops.iou_matrix(_dev(np.concatenate(b0), np.float64), _dev(np.concatenate(b1), np.float64),
                                        _dev(np.array(off0), np.int64), _dev(np.array(off1), np.int64))
        flat, out_off = flat.cpu().numpy(), out_off.cpu().numpy()

    # ---- pass 3 (host): the assignment per scene, the JSON
    annotations = []
    g = 0
    for sc in scenes:
        if sc["match"]:
            n0, n1 = len(sc["annos"]), len(sc["bbox2d"])
            iou = flat[out_off[g]:out_off[g + 1]].reshape(n0, n1)
            for i, j, _ in _match(iou):
                sc["annos"][i]["bbox2D_tight"] = sc["bbox2d"][j]
            g += 1
        else:
            for anno in sc["annos"]:
                anno["bbox2D_tight"] = anno["bbox2D_trunc"]
        annotations.extend(sc["annos"])

    output = {"info": {"id": dataset_id, "source": "COCO", "name": f"COCO {'Validation' if split == 'val' else 'Train'}",
                       "split": split.capitalize(), "version": "0.1", "url": "https://cocodataset.org/#home"},
              "categories": COCO_CATEGORIES, "images": images, "annotations": annotations}
    os.makedirs(os.path.dirname(output_path) if os.path.dirname(output_path) else ".", exist_ok=True)
    with open(output_path, "w") as f:
        json.dump(output, f)
    print(f"Saved {len(images)} images, {len(annotations)} annotations to {output_path}")


def main(argv=None):
    """Same command line as the reference script (:314-326)."""
    cli = argparse.ArgumentParser(description="Combine COCO 3D bbox results into Omni3D format")
    cli.add_argument("--split", default="val", choices=["train", "val"], type=str, help="Dataset split")
    cli.add_argument("--results_dir", default="../experimental_results/COCO", type=str, help="Results directory")
    cli.add_argument("--output", default=None, type=str, help="Output JSON path")
    cli.add_argument("--bbox_file", default="3dbbox.json", type=str, help="3D bbox JSON filename")
    opts = cli.parse_args(argv)
    target = opts.output if opts.output is not None else os.path.join(opts.results_dir, f"COCO3D_{opts.split}.json")
    combine_coco_results(opts.results_dir, opts.split, target, opts.bbox_file)


if __name__ == "__main__":
    main()
