"""Footprints whose hull edges tie exactly or almost exactly in bounding-rectangle area: axis-aligned rectangles
(filled lattices) at several rotations and regular n-gons.  The reference's `_estimate_yaw_convex_hull`
(`src/util_3dbox.py:189-224`) keeps the FIRST strict minimum in SciPy / Qhull's vertex order, whose starting
vertex is a Qhull implementation detail; this repository's hull starts at the lexicographically smallest point.
The golden file made from these cases records which of them the reference resolves differently."""
import math

import numpy as np


def cases():
    out = {}
    rng = np.random.RandomState(0)
    for name, (w, h) in {"rect2x1": (2.0, 1.0), "square": (1.0, 1.0)}.items():
        for ang in (0, 15, 30, 45, 60, 90, 137):
            xs, zs = np.meshgrid(np.linspace(-w / 2, w / 2, 9), np.linspace(-h / 2, h / 2, 5))
            p = np.stack([xs.ravel(), zs.ravel()], 1)
            a = math.radians(ang)
            R = np.array([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
            q = p @ R.T + np.array([0.3, 4.0])
            y = rng.uniform(-0.5, 0.5, len(q))
            out[f"{name}_rot{ang}"] = np.stack([q[:, 0], y, q[:, 1]], 1)
    for n in (3, 4, 5, 6, 8, 12):
        for ang in (0, 10):
            t = np.arange(n) * 2 * math.pi / n + math.radians(ang)
            ring = np.stack([np.cos(t), np.sin(t)], 1)
            p = np.concatenate([ring, ring * 0.5, [[0, 0]]]) + np.array([1.0, 5.0])
            y = rng.uniform(-0.5, 0.5, len(p))
            out[f"ngon{n}_rot{ang}"] = np.stack([p[:, 0], y, p[:, 1]], 1)
    return out


def edge_search(pc):
    """``(areas, yaws)`` of the reference's hull-edge search for every hull edge (this repository's vertex order)."""
    from oracle import la3d_oracle as orc
    xz = np.asarray(pc)[:, [0, 2]]
    hull = xz[orc.convex_hull_ccw(xz)]
    areas, yaws = [], []
    for i in range(len(hull)):
        e = hull[(i + 1) % len(hull)] - hull[i]
        yaw = np.arctan2(e[1], e[0])
        c, s = np.cos(yaw), np.sin(yaw)
        rot = (np.array([[c, -s], [s, c]]) @ xz.T).T
        areas.append((rot[:, 0].max() - rot[:, 0].min()) * (rot[:, 1].max() - rot[:, 1].min()))
        yaws.append(yaw)
    return np.array(areas), np.array(yaws)


def edge_areas(pc):
    return edge_search(pc)[0]
