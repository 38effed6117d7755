"""Drop-in for the reference's ``src/util_3dbox.py``: same names, same signatures, same
return types and error behaviour - with the box arithmetic running on the B200.

Put this directory in front of ``sys.path`` (``labelany3d_b200.dropin_path()``) and the
pipeline's ``from util_3dbox import save_3d_with_ground_alignment_bbox``
(``src/batch_scripts/whole.py:16`` of the reference) resolves here unchanged.

``estimate_bbox`` is ``la3d_fit_points`` (``include/la3d.h``) on one point set; the small
matrix helpers are host-side NumPy one-liners exactly as in the reference API (they are
not on the hot path: the kernel carries its own copies of that arithmetic).
"""

from __future__ import annotations

import json
import math
import os

import numpy as np
import torch

from labelany3d_b200 import ops as _ops
from labelany3d_b200 import records as _rec


# ---------------------------------------------------------------- basic geometry (API parity)
def normalize(v):
    """``src/util_3dbox.py:20-25``."""
    n = np.linalg.norm(v)
    return v if n == 0 else v / n


def rotate_y(yaw):
    """``src/util_3dbox.py:28-34``."""
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rotation_matrix_from_vectors(vec1, vec2):
    """``src/util_3dbox.py:37-55`` (Rodrigues; undefined for parallel inputs, as there)."""
    a, b = normalize(vec1), normalize(vec2)
    axis = np.cross(a, b)
    cos_theta = np.dot(a, b)
    S = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + S + np.dot(S, S) * (1 - cos_theta) / (np.linalg.norm(axis) ** 2)


def point_to_plane_distance(plane, x, y, z):
    """``src/util_3dbox.py:58-64``."""
    a, b, c, d = np.array(plane)
    return abs(a * x + b * y + c * z + d) / np.sqrt(a ** 2 + b ** 2 + c ** 2)


def convert_box_vertices(center_x, center_y, center_z, l, w, h, yaw):
    """``src/util_3dbox.py:71-103``: 8 corners, order (-,-,-),(+,-,-),(+,+,-),(-,+,-),(-,-,+),..."""
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1]) * (l / 2)
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1]) * (w / 2)
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1]) * (h / 2)
    local = np.stack([sx, sy, sz], axis=1)
    rot = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
    return np.dot(local, rot.T) + np.array([center_x, center_y, center_z])


# ---------------------------------------------------------------- box fit on the GPU
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("labelany3d_b200 needs a CUDA device: the box fit has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def _points64(in_pc):
    if isinstance(in_pc, torch.Tensor):
        in_pc = in_pc.detach().cpu().numpy()
    pc = np.asarray(in_pc)
    if pc.ndim != 2 or pc.shape[1] != 3:
        raise ValueError(f"expected points of shape [N,3], got {pc.shape}")
    return np.ascontiguousarray(pc, dtype=np.float64)


def _fit_one(pc, ground_equ, method, yaw_steps=0, K=None, sample_idx=None):
    """One point set -> one unpacked record (device round trip)."""
    dev = _device()
    g = None
    if ground_equ is not None:
        g = torch.as_tensor(np.asarray(ground_equ, dtype=np.float64)[:3].reshape(1, 3).copy(), device=dev)
    idx = None
    if sample_idx is not None:
        idx = torch.as_tensor(np.asarray(sample_idx, dtype=np.int32).reshape(1, -1), device=dev)
    Kd = None if K is None else torch.as_tensor(np.asarray(K, dtype=np.float64).reshape(1, 3, 3).copy(), device=dev)
    rec = _ops.fit_points(torch.as_tensor(pc, device=dev),
                          torch.tensor([0, pc.shape[0]], dtype=torch.int64, device=dev),
                          idx, Kd, g, method, yaw_steps)
    return _rec.unpack(rec[0].cpu().numpy())


def estimate_bbox(in_pc, cat_name=None, ground_equ=None, method='pca', yaw_steps=None, verbose=True):
    """Oriented box of a point cloud; ``src/util_3dbox.py:106-178``.

    Returns ``(vertices[8,3], center_cam[3], [dz, dy, dx], R_cam[3,3])`` as NumPy/float64 like
    the reference.  More than 500 points are subsampled with ``np.random.randint(0, N, 500)``
    drawn from NumPy's global legacy generator exactly once, as the reference does.  The same
    ``ValueError``s are raised (no valid points, unknown method, and the two that surface from
    scikit-learn in the reference: infinite footprint, single sample).  ``method='sweep'`` with
    ``yaw_steps=K`` is an addition (uniform yaw sweep).  ``cat_name`` is unused, as there.
    """
    pc = _points64(in_pc)
    sample_idx = None
    if pc.shape[0] > _rec.SUBSAMPLE:
        sample_idx = np.random.randint(0, pc.shape[0], _rec.SUBSAMPLE)
    r = _fit_one(pc, ground_equ, method, int(yaw_steps or 0), sample_idx=sample_idx)
    err = _rec.status_error(r["status"], method, r["n_valid"])
    if err is not None:
        raise err
    if r["hull_fallback"]:          # the reference prints Qhull's own message here (:222-224); the event is what matters
        print("ConvexHull failed: degenerate footprint (coincident or collinear points), falling back to PCA")
    dz, dy, dx = (np.float64(v) for v in r["dimensions"])
    if verbose:
        print(f"[{method}] dx={dx:.3f}, dy={dy:.3f}, dz={dz:.3f}")
    return r["bbox3D_cam"], r["center_cam"], [dz, dy, dx], r["R_cam"]


def _yaw_only(rotated_pc, method):
    pc = _points64(rotated_pc)
    if pc.shape[0] > _rec.SUBSAMPLE:
        raise ValueError("the yaw helpers take at most 500 points (estimate_bbox subsamples before calling them)")
    r = _fit_one(pc, None, method)
    err = _rec.status_error(r["status"], method, r["n_valid"])
    if err is not None:
        raise err
    if r["hull_fallback"]:
        print("ConvexHull failed: degenerate footprint (coincident or collinear points), falling back to PCA")
    return np.float64(r["yaw"])


def _estimate_yaw_pca(rotated_pc):
    """``src/util_3dbox.py:181-186``."""
    return _yaw_only(rotated_pc, "pca")


def _estimate_yaw_convex_hull(rotated_pc):
    """``src/util_3dbox.py:189-224`` (falls back to PCA when the hull is degenerate)."""
    return _yaw_only(rotated_pc, "convex_hull")


# ---------------------------------------------------------------- scene driver
def _load_object_points(mesh_path):
    """500 surface samples of a reconstructed object (``src/util_3dbox.py:256-270``); ``None`` = skip."""
    import trimesh   # imported lazily: only this driver needs it
    mesh = trimesh.load(mesh_path)
    if isinstance(mesh, trimesh.Scene):
        mesh = mesh.dump()[0]
    if mesh.is_empty or mesh.area == 0 or len(mesh.faces) == 0:
        print(f"Invalid mesh at {mesh_path}, skipping.")
        return None
    return np.array(trimesh.points.PointCloud(mesh.sample(500)).vertices)


def save_3d_with_ground_alignment_bbox(scene_dir, bbox_method='pca'):
    """Boxes for every ``reconstruction/*.glb`` of a scene -> ``3dbbox_ground.json``
    (``src/util_3dbox.py:231-294``); returns the list of dictionaries it wrote.

    All objects of the scene are fitted in ONE ``la3d_fit_points`` launch; objects whose fit
    fails are reported and skipped like the reference's per-object ``try/except``.
    """
    recons_dir = os.path.join(scene_dir, "reconstruction")
    objs = [f for f in os.listdir(recons_dir)
            if f not in ("full_scene.glb", "background.ply") and f.endswith(".glb")]
    metas, clouds, grounds = [], [], []
    for obj in objs:
        obj_id, rest = obj.split("_", 1)
        category = rest.split(".", 1)[0]
        upright = np.load(os.path.join(recons_dir, f"{obj.split('.', 1)[0]}_canonical_upright.npy"))
        pts = _load_object_points(os.path.join(recons_dir, obj))
        if pts is None:
            continue
        metas.append((obj, obj_id, category))
        clouds.append(_points64(pts))
        grounds.append(np.asarray(upright, dtype=np.float64).reshape(-1)[:3])

    bbox_list = []
    if clouds:
        dev = _device()
        offsets = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
        idx = np.zeros((len(clouds), _rec.SUBSAMPLE), dtype=np.int32)
        for j, c in enumerate(clouds):
            if len(c) > _rec.SUBSAMPLE:
                idx[j] = np.random.randint(0, len(c), _rec.SUBSAMPLE)
        rec = _ops.fit_points(torch.as_tensor(np.concatenate(clouds), device=dev), torch.as_tensor(offsets, device=dev),
                              torch.as_tensor(idx, device=dev), None, torch.as_tensor(np.stack(grounds), device=dev),
                              bbox_method).cpu().numpy()
        for (obj, obj_id, category), row in zip(metas, rec):
            r = _rec.unpack(row)
            err = _rec.status_error(r["status"], bbox_method, r["n_valid"])
            if err is not None:
                print(f"Error estimating bbox for {obj}: {err}")
                continue
            dz, dy, dx = r["dimensions"]
            print(f"[{bbox_method}] dx={dx:.3f}, dy={dy:.3f}, dz={dz:.3f}")
            bbox_list.append({
                "obj_id": obj_id,
                "category_name": category,
                "center_cam": r["center_cam"].tolist(),
                "R_cam": r["R_cam"].tolist(),
                "dimensions": [float(dz), float(dy), float(dx)],
                "bbox3D_cam": r["bbox3D_cam"].tolist(),
            })
    with open(os.path.join(scene_dir, "3dbbox_ground.json"), "w") as f:
        json.dump(bbox_list, f)
    return bbox_list
