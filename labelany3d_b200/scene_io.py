"""On-disk formats either side of the box-fitting path (scope-table row f4): the files the reference's
stages exchange through a scene directory, so that the batched GPU path can be fed from, and feed, the
unmodified rest of the pipeline.

====================  =====================================================  ==============================
file                  written by (reference)                                 here
====================  =====================================================  ==============================
``depth_map.npy``     ``src/batch_scripts/depth.py:156``  float32 ``[H,W]``  ``save_depth_stage`` / ``load_scene``
``cam_params.json``   ``src/batch_scripts/depth.py:158-167``                 ``save_depth_stage`` / ``load_scene``
                      ``{"K", "c2w", "W", "H"}``
``3dbbox.json`` /     ``src/util_3dbox.py:283-292``  list of                 ``records_to_bbox_list`` /
``3dbbox_ground.json``  ``{obj_id, category_name, center_cam, R_cam,         ``save_bbox_json``
                      dimensions, bbox3D_cam}``                              (read by ``draw_cube``, ``combine_results``)
====================  =====================================================  ==============================

``fit_scene_dirs`` chains them: scene directories + their instance masks -> ONE batched ``fit_boxes``
launch sequence -> one ``3dbbox*.json`` per scene.  Host code only; the arithmetic is ``ops``.
"""

from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import ops as _ops
from . import records as _rec


def save_depth_stage(out_dir, depth_map, K, width, height, c2w=None):
    """``depth_map.npy`` + ``cam_params.json`` exactly as the depth stage leaves them
    (``src/batch_scripts/depth.py:156-167``)."""
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, "depth_map.npy"), depth_map)
    pose = np.eye(4) if c2w is None else np.asarray(c2w)
    cam_params = {"K": np.asarray(K).tolist(), "c2w": pose.tolist(), "W": width, "H": height}
    with open(os.path.join(out_dir, "cam_params.json"), "w") as fp:
        json.dump(cam_params, fp)


def load_scene(scene_dir):
    """``(depth_map float32 [H,W], K float64 [3,3], W, H)`` of a scene directory."""
    depth = np.load(os.path.join(scene_dir, "depth_map.npy"))
    with open(os.path.join(scene_dir, "cam_params.json")) as fp:
        cam = json.load(fp)
    return np.ascontiguousarray(depth, dtype=np.float32), np.array(cam["K"], dtype=np.float64), int(cam["W"]), int(cam["H"])


def records_to_bbox_list(records, obj_ids, categories, method="pca", on_error=print):
    """Packed records ``[n,64]`` -> the reference's list of box dictionaries
    (``src/util_3dbox.py:283-288``); boxes whose status is an error are reported through
    ``on_error`` and skipped, like the reference's per-object ``try/except`` (``:279-281``)."""
    out = []
    for row, obj_id, cat in zip(np.asarray(records, dtype=np.float64), obj_ids, categories):
        r = _rec.unpack(row)
        err = _rec.status_error(r["status"], method, r["n_valid"])
        if r["status"] < 0:
            continue                                  # padding slot of a sharded gather
        if err is not None:
            if on_error is not None:
                on_error(f"Error estimating bbox for {obj_id}_{cat}: {err}")
            continue
        dz, dy, dx = r["dimensions"]
        out.append({"obj_id": obj_id, "category_name": cat, "center_cam": r["center_cam"].tolist(),
                    "R_cam": r["R_cam"].tolist(), "dimensions": [float(dz), float(dy), float(dx)],
                    "bbox3D_cam": r["bbox3D_cam"].tolist()})
    return out


def save_bbox_json(scene_dir, bbox_list, is_ground=False):
    """``3dbbox_ground.json`` / ``3dbbox.json`` (``src/util_3dbox.py:291-292``; the names ``draw_cube`` reads)."""
    with open(os.path.join(scene_dir, "3dbbox_ground.json" if is_ground else "3dbbox.json"), "w") as f:
        json.dump(bbox_list, f)


def fit_scene_dirs(scene_dirs, masks, categories, grounds=None, method="pca", yaw_steps=0, seed=0, device="cuda"):
    """Boxes for many scene directories in one batch.

    ``masks[s]``: bool ``[I_s,H,W]`` instance masks of scene ``s`` (the stack of ``src/util.py:382``);
    ``categories[s]``: ``I_s`` category names; ``grounds[s]``: ``[I_s,3]`` ground normals or ``None``.
    Scenes must share ``H x W``; scenes with fewer instances are padded with empty masks (their boxes
    report "no valid points" and are dropped).  Writes ``3dbbox_ground.json`` (or ``3dbbox.json`` without
    ground) into every scene directory and returns the per-scene lists.
    """
    dev = torch.device(device)
    loaded = [load_scene(d) for d in scene_dirs]
    H, W = loaded[0][0].shape
    if any(l[0].shape != (H, W) for l in loaded):
        raise ValueError("fit_scene_dirs: all scenes of a batch must share one image size")
    B, I = len(scene_dirs), max(max(len(m) for m in masks), 1)
    depth = torch.as_tensor(np.stack([l[0] for l in loaded]), device=dev)
    K = torch.as_tensor(np.stack([l[1] for l in loaded]), device=dev)
    stack = np.zeros((B, I, H, W), dtype=bool)
    g = None if grounds is None else np.tile(np.array([0.0, -1.0, 0.0]), (B, I, 1))
    for s, m in enumerate(masks):
        if len(m):
            stack[s, :len(m)] = np.asarray(m) != 0
            if grounds is not None:
                g[s, :len(m)] = np.asarray(grounds[s], dtype=np.float64)[:, :3]
    rec = _ops.fit_boxes(depth, K, torch.as_tensor(stack, device=dev), None if g is None else torch.as_tensor(g, device=dev),
                         method, yaw_steps, seed=seed).cpu().numpy()
    out = []
    for s, d in enumerate(scene_dirs):
        n = len(masks[s])
        boxes = records_to_bbox_list(rec[s, :n], [str(i) for i in range(n)], list(categories[s]), method)
        save_bbox_json(d, boxes, is_ground=grounds is not None)
        out.append(boxes)
    return out
