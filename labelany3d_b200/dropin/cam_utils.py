"""Drop-in for the reference's ``src/cam_utils.py`` (camera helpers) plus a re-export of
``depth_to_points``, which BASELINE.json's north_star places here.

``length`` / ``safe_normalize`` / ``look_at`` / ``orbit_camera`` are a handful of NumPy lines
in the reference, imported by nothing on the hot path; they stay NumPy here.  The reference's
torch branch of ``length`` calls an undefined ``dot`` (``src/cam_utils.py:8``); here it works.
"""

from __future__ import annotations

import numpy as np
import torch

from util import depth_to_points  # noqa: F401  (the B200 back-projection)


def length(x, eps=1e-20):
    if isinstance(x, np.ndarray):
        return np.sqrt(np.maximum(np.sum(x * x, axis=-1, keepdims=True), eps))
    return torch.sqrt(torch.clamp(torch.sum(x * x, dim=-1, keepdim=True), min=eps))


def safe_normalize(x, eps=1e-20):
    return x / length(x, eps)


def look_at(campos, target, opengl=True):
    """Rotation ``[.., 3, 3]`` whose columns are right / up / forward (``src/cam_utils.py:14-32``)."""
    up = np.array([0, 1, 0], dtype=np.float32)
    if opengl:      # camera forward aligns with +z
        forward = safe_normalize(campos - target)
        right = safe_normalize(np.cross(up, forward))
        up = safe_normalize(np.cross(forward, right))
    else:           # camera forward aligns with -z
        forward = safe_normalize(target - campos)
        right = safe_normalize(np.cross(forward, up))
        up = safe_normalize(np.cross(right, forward))
    return np.stack([right, up, forward], axis=1)


def orbit_camera(elevation, azimuth, radius=1, is_degree=True, target=None, opengl=True):
    """Elevation / azimuth -> 4x4 camera-to-world pose (``src/cam_utils.py:35-52``)."""
    if is_degree:
        elevation, azimuth = np.deg2rad(elevation), np.deg2rad(azimuth)
    pos = np.array([radius * np.cos(elevation) * np.sin(azimuth),
                    -radius * np.sin(elevation),
                    radius * np.cos(elevation) * np.cos(azimuth)])
    if target is None:
        target = np.zeros([3], dtype=np.float32)
    campos = pos + target
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = look_at(campos, target, opengl)
    T[:3, 3] = campos
    return T
