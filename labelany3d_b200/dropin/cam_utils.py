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
    """``sqrt(max(sum(x*x, -1), eps))`` with the last axis kept (``src/cam_utils.py:4-8``)."""
    sq = x * x
    if isinstance(x, np.ndarray):
        return np.sqrt(np.maximum(sq.sum(axis=-1, keepdims=True), eps))
    return sq.sum(dim=-1, keepdim=True).clamp(min=eps).sqrt()


def safe_normalize(x, eps=1e-20):
    return x / length(x, eps)


def look_at(campos, target, opengl=True):
    """Rotation ``[.., 3, 3]`` whose columns are right / up / forward (``src/cam_utils.py:14-32``).
    ``opengl``: the camera looks down -z (forward = campos - target), else down +z."""
    world_up = np.array([0, 1, 0], dtype=np.float32)
    sign = 1.0 if opengl else -1.0
    forward = safe_normalize(sign * (campos - target))
    # right = up x forward (OpenGL) or forward x up: the same vector once forward flips sign
    right = safe_normalize(np.cross(world_up, forward) if opengl else np.cross(forward, world_up))
    up = safe_normalize(np.cross(forward, right) if opengl else np.cross(right, forward))
    return np.stack([right, up, forward], axis=1)


def orbit_camera(elevation, azimuth, radius=1, is_degree=True, target=None, opengl=True):
    """Elevation / azimuth -> 4x4 camera-to-world pose (``src/cam_utils.py:35-52``): elevation runs
    from +y to -y over (-90, 90), azimuth from +z to +x over (0, 90)."""
    el, az = (np.deg2rad(elevation), np.deg2rad(azimuth)) if is_degree else (elevation, azimuth)
    offset = np.array([radius * np.cos(el) * np.sin(az), -radius * np.sin(el), radius * np.cos(el) * np.cos(az)])
    centre = np.zeros([3], dtype=np.float32) if target is None else target
    eye = offset + centre
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = look_at(eye, centre, opengl)
    pose[:3, 3] = eye
    return pose
