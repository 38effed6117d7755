"""CPU oracle for the LabelAny3D 3D-box-fitting hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's algorithm.  It is the
checker the CUDA path is compared against; it is never the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  Nothing under
``labelany3d_b200/`` imports it, and the product path has no CPU fallback.

Parity status: PINNED.  Every function below is checked against the unmodified
reference functions (imported live from ``/root/reference/src`` where that tree
exists) by ``tests/golden/make_golden.py``, which also writes the committed
golden vectors ``tests/golden/*.npz`` that the CPU and GPU test-suites replay.
The reference itself ships no tests or golden vectors, so
the live-imported reference is the only pin there is.

Reference lines each function follows (paths relative to the reference root):

==============================  ===============================================
oracle function                 reference
==============================  ===============================================
``depth_to_points``             ``src/util.py:52-75``
``masked_points``               NumPy idiom ``pts[mask]`` (``src/util.py:480-481``)
``unit`` / ``yaw_matrix`` /     ``src/util_3dbox.py:20-25, 28-34, 37-55``
``rotation_between``
``box_corners``                 ``src/util_3dbox.py:71-103``
``estimate_bbox``               ``src/util_3dbox.py:106-178``
``yaw_from_pca``                ``src/util_3dbox.py:181-186`` (+ scikit-learn
                                ``PCA(2)``, pinned 1.5.0, run here on 1.9.0)
``yaw_from_hull``               ``src/util_3dbox.py:189-224`` (+ SciPy/Qhull
                                ``ConvexHull``, pinned 1.14.0, here 1.18.1)
``yaw_from_sweep``              NEW feature, no reference code; SURVEY.md
                                section 8 row a7 is its specification
``project_to_2d`` /             ``src/util.py:227-229``,
``box2d_from_corners``          ``src/tools/combine_results.py:105-108, 238-252``
``LegacyMT19937`` /             NumPy legacy ``RandomState.randint`` stream
``legacy_randint``              (``src/util_3dbox.py:123-125`` call site;
                                numpy pinned 1.24.4, here 2.3.5)
``fit_boxes``                   composition of SURVEY.md section 3.4
==============================  ===============================================

Third-party arithmetic that is not part of the reference tree (scikit-learn PCA,
SciPy ConvexHull, NumPy's MT19937) is used here the same way the reference uses
it, and *also* restated in closed form (``yaw_from_pca(..., impl="closed")``,
``convex_hull_ccw``, ``LegacyMT19937``) because those closed forms are what the
CUDA kernels implement.  Tests assert the two agree.
"""

from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# Packed per-box record (SURVEY.md section 8 row e).  64 scalars.
# --------------------------------------------------------------------------
REC = 64
O_VERT = 0        # 8 x 3 box corners in the camera frame
O_CENTER = 24     # center_cam (3)
O_DIM = 27        # [dz, dy, dx]
O_RCAM = 30       # R_cam row-major (9)
O_YAW = 39
O_NVALID = 40     # points that survived the NaN filter
O_STATUS = 41     # 0 ok / 1 no valid points / 2 PCA undefined (<2 points) / 3 unknown method / 4 inf in footprint
O_UV = 42         # 8 x 2 projected corners
O_BOX2D = 58      # [min u, min v, max u, max v]
O_NMASK = 62      # pixels set in the mask (or points given)
O_PAD = 63

ST_OK, ST_NO_VALID, ST_PCA_UNDEFINED, ST_BAD_METHOD, ST_NONFINITE = 0, 1, 2, 3, 4

# Messages of the exceptions the reference path raises (the last two come from
# scikit-learn's input validation inside ``PCA.fit``).
MSG_NO_VALID = "No valid points after removing NaN values"
MSG_NONFINITE = "Input X contains infinity or a value too large for dtype('float64')."


def status_of_exception(exc):
    text = str(exc)
    if "No valid points" in text:
        return ST_NO_VALID
    if "contains infinity" in text:
        return ST_NONFINITE
    if "n_components" in text:
        return ST_PCA_UNDEFINED
    if "Unknown method" in text:
        return ST_BAD_METHOD
    raise exc

SUBSAMPLE = 500   # src/util_3dbox.py:123-124


# --------------------------------------------------------------------------
# a1  depth -> camera-space points                     src/util.py:52-75
# --------------------------------------------------------------------------
def depth_to_points(depth, K=None, R=None, t=None):
    """Back-project ``depth[bs,H,W]`` with intrinsics ``K``; returns element 0.

    Integer pixel coordinates (no half-pixel offset), homogeneous coordinate
    in float32, the scaled inverse intrinsics and the products in float64,
    optional rigid transform afterwards.  The output is ``[H,W,3]`` float64.
    """
    Kinv = np.linalg.inv(K)
    rot = np.eye(3) if R is None else R
    trans = np.zeros(3) if t is None else t
    H, W = depth.shape[1:3]
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    homog = np.stack((uu, vv, np.ones_like(uu)), axis=-1).astype(np.float32)
    scaled = depth[:, :, :, None, None] * Kinv[None, None, None, :, :]
    cam = scaled @ homog[None, :, :, :, None]
    world = rot[None, None, None, :, :] @ cam + trans[None, None, None, :, None]
    return world[:, :, :, :3, 0][0]


def depth_to_points_closed(depth2d, Kinv, R=None, t=None):
    """Same numbers as :func:`depth_to_points` written out per component.

    ``x_i = ((d*Kinv[i,0])*u + (d*Kinv[i,1])*v) + (d*Kinv[i,2])*1`` in float64
    with no fused multiply-add, then ``R @ x + t`` summed left to right.  This
    is the exact arithmetic the CUDA lift kernel performs.
    """
    H, W = depth2d.shape
    d = depth2d.astype(np.float64)
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    cam = [((d * Kinv[i, 0]) * u + (d * Kinv[i, 1]) * v) + (d * Kinv[i, 2]) * 1.0 for i in range(3)]
    if R is None and t is None:
        return np.stack(cam, -1)
    rot = np.eye(3) if R is None else R
    trans = np.zeros(3) if t is None else t
    out = [((rot[i, 0] * cam[0] + rot[i, 1] * cam[1]) + rot[i, 2] * cam[2]) + trans[i] for i in range(3)]
    return np.stack(out, -1)


# --------------------------------------------------------------------------
# a2  per-instance masked gather
# --------------------------------------------------------------------------
def masked_points(points_hw3, mask_hw):
    """Row-major compaction of the pixels where ``mask`` is set -> ``[N,3]``."""
    return points_hw3[np.asarray(mask_hw).astype(bool)]


def mask_counts(masks):
    """``masks[...,H,W]`` -> number of set pixels per plane (int64)."""
    m = np.asarray(masks) != 0
    return m.reshape(m.shape[:-2] + (-1,)).sum(-1)


def mask_select(mask_hw, ranks):
    """Flat pixel index (v*W+u) of the r-th set pixel in row-major order."""
    flat = np.flatnonzero(np.asarray(mask_hw).reshape(-1) != 0)
    return flat[np.asarray(ranks)]


# --------------------------------------------------------------------------
# a4  small geometry helpers               src/util_3dbox.py:20-55, 71-103
# --------------------------------------------------------------------------
def unit(v):
    n = np.linalg.norm(v)
    return v if n == 0 else v / n


def yaw_matrix(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rotation_between(a, b):
    """Rodrigues rotation taking direction ``a`` to ``b`` (0/0 when parallel)."""
    a = unit(a)
    b = unit(b)
    ax = np.cross(a, b)
    cosang = np.dot(a, b)
    S = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + S + np.dot(S, S) * (1 - cosang) / (np.linalg.norm(ax) ** 2)


def point_to_plane_distance(plane, x, y, z):
    a, b, c, d = np.array(plane)
    return abs(a * x + b * y + c * z + d) / np.sqrt(a ** 2 + b ** 2 + c ** 2)


_CORNER_SIGNS = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def box_corners(cx, cy, cz, l, w, h, yaw):
    """Eight corners, order (-,-,-),(+,-,-),(+,+,-),(-,+,-),(-,-,+),(+,-,+),(+,+,+),(-,+,+)."""
    half = np.array([l / 2, w / 2, h / 2])
    local = _CORNER_SIGNS * half
    rot = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
    return np.dot(local, rot.T) + np.array([cx, cy, cz])


# --------------------------------------------------------------------------
# a5  yaw from PCA of the XZ footprint          src/util_3dbox.py:181-186
# --------------------------------------------------------------------------
def yaw_from_pca(pc, impl="sklearn"):
    """Heading of the first principal axis of ``pc[:, [0, 2]]``.

    ``impl="sklearn"`` calls scikit-learn exactly as the reference does.
    ``impl="closed"`` is the closed form the CUDA kernel evaluates: with
    ``C = (X^T X - n mu mu^T)/(n-1) = [[a,b],[b,c]]``, ``theta = atan2(2b, a-c)/2``,
    ``(vx,vz) = (cos theta, sin theta)`` and scikit-learn's sign rule (the
    entry of larger magnitude is made positive; ties go to vx).
    """
    xz = pc[:, [0, 2]]
    if impl == "sklearn":
        from sklearn.decomposition import PCA
        comp = PCA(2).fit(xz).components_[0, :]
        return np.arctan2(comp[1], comp[0])
    n = xz.shape[0]
    if not np.isfinite(xz).all():
        raise ValueError(MSG_NONFINITE)
    if n < 2:
        raise ValueError(
            f"n_components=2 must be between 0 and min(n_samples, n_features)={min(n, 2)} "
            "with svd_solver='full'")
    x = xz[:, 0].astype(np.float64)
    z = xz[:, 1].astype(np.float64)
    mx, mz = x.sum() / n, z.sum() / n
    a = ((x * x).sum() - n * mx * mx) / (n - 1)
    b = ((x * z).sum() - n * mx * mz) / (n - 1)
    c = ((z * z).sum() - n * mz * mz) / (n - 1)
    theta = 0.5 * math.atan2(2.0 * b, a - c)
    vx, vz = math.cos(theta), math.sin(theta)
    if abs(vx) >= abs(vz):
        if vx < 0:
            vx, vz = -vx, -vz
    elif vz < 0:
        vx, vz = -vx, -vz
    return math.atan2(vz, vx)


# --------------------------------------------------------------------------
# a6  yaw from convex-hull edges                src/util_3dbox.py:189-224
# --------------------------------------------------------------------------
def convex_hull_ccw(xz):
    """Strict convex hull of 2-D points, counter-clockwise, as point indices.

    Andrew's monotone chain; collinear points are dropped.  Same vertex *set*
    and cyclic order as SciPy/Qhull in 2-D (Qhull's starting vertex is an
    implementation detail; see :func:`yaw_from_hull`).
    """
    pts = np.asarray(xz, dtype=np.float64)
    if not np.isfinite(pts).all():
        raise ValueError("hull input is not finite")      # Qhull refuses too (QH6214 et al.)
    order = np.lexsort((pts[:, 1], pts[:, 0]))
    # drop exact duplicates
    uniq = [order[0]]
    for i in order[1:]:
        if pts[i, 0] != pts[uniq[-1], 0] or pts[i, 1] != pts[uniq[-1], 1]:
            uniq.append(i)
    if len(uniq) < 3:
        raise ValueError("hull needs 3 distinct points")

    def cross(o, a, b):
        return (pts[a, 0] - pts[o, 0]) * (pts[b, 1] - pts[o, 1]) - (pts[a, 1] - pts[o, 1]) * (pts[b, 0] - pts[o, 0])

    lower = []
    for i in uniq:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], i) <= 0:
            lower.pop()
        lower.append(i)
    upper = []
    for i in reversed(uniq):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], i) <= 0:
            upper.pop()
        upper.append(i)
    hull = lower[:-1] + upper[:-1]
    if len(hull) < 3:
        raise ValueError("points are collinear")
    return np.array(hull)


def _hull_search(xz, hull_xy):
    """Loop of src/util_3dbox.py:202-218 over hull vertices in the given order."""
    best_area = float("inf")
    best_yaw = 0
    m = len(hull_xy)
    for i in range(m):
        e = hull_xy[(i + 1) % m] - hull_xy[i]
        yaw = np.arctan2(e[1], e[0])
        c, s = np.cos(yaw), np.sin(yaw)
        # NOTE the reference rotates the footprint by +yaw here but builds the
        # box with yaw_matrix(yaw), which rotates XZ by -yaw.  Kept as is.
        rot = (np.array([[c, -s], [s, c]]) @ xz.T).T
        area = (rot[:, 0].max() - rot[:, 0].min()) * (rot[:, 1].max() - rot[:, 1].min())
        if area < best_area:
            best_area = area
            best_yaw = yaw
    return best_yaw


def yaw_from_hull(pc, impl="scipy", verbose=False, info=None):
    """Hull-edge search.  Any hull failure falls back to PCA (``:222-224``); ``info["fallback"]`` says so."""
    xz = pc[:, [0, 2]]
    try:
        if impl == "scipy":
            from scipy.spatial import ConvexHull
            idx = ConvexHull(xz).vertices
        else:
            idx = convex_hull_ccw(xz)
        return _hull_search(xz, xz[idx])
    except Exception as exc:  # noqa: BLE001 - the reference catches everything
        if verbose:
            print(f"ConvexHull failed: {exc}, falling back to PCA")
        if info is not None:
            info["fallback"] = True
        return yaw_from_pca(pc, impl="sklearn" if impl == "scipy" else "closed")


# --------------------------------------------------------------------------
# a7  uniform yaw sweep (NEW)
# --------------------------------------------------------------------------
def yaw_from_sweep(pc, steps):
    """First strict minimum of ``dx*dz`` over ``yaw_k = k*(pi/2)/steps``."""
    best_area = float("inf")
    best_yaw = 0.0
    for k in range(int(steps)):
        yaw = k * (np.pi / 2) / steps
        r = yaw_matrix(yaw) @ pc.T
        area = (r[0].max() - r[0].min()) * (r[2].max() - r[2].min())
        if area < best_area:
            best_area = area
            best_yaw = yaw
    return best_yaw


# --------------------------------------------------------------------------
# a3  oriented box from a point set            src/util_3dbox.py:106-178
# --------------------------------------------------------------------------
def ground_rotation(ground):
    """Rg of ``:128-134``: align (0,-1,0) with the (sign-fixed) ground normal."""
    if ground is None:
        return np.eye(3)
    g = np.asarray(ground)
    if np.dot([0, -1, 0], g[:3]) <= 0:
        g = -g
    return rotation_between([0, -1, 0], g[:3])


def fit_details(in_pc, ground_equ=None, method="pca", yaw_steps=None, rng=None, impl="library",
                verbose=False, sample_idx=None):
    """Body of ``estimate_bbox`` returning every intermediate the tests look at."""
    pc = np.asarray(in_pc)
    if pc.shape[0] > SUBSAMPLE:
        if sample_idx is None:
            sample_idx = (np.random if rng is None else rng).randint(0, pc.shape[0], SUBSAMPLE)
        pc = pc[sample_idx]
    Rg = ground_rotation(ground_equ)
    aligned = np.dot(pc, Rg)
    aligned = aligned[~np.isnan(aligned).any(axis=1)]
    if len(aligned) == 0:
        raise ValueError(MSG_NO_VALID)

    lib = impl == "library"
    hull_info = {}
    if method == "convex_hull":
        yaw = yaw_from_hull(aligned, impl="scipy" if lib else "closed", verbose=verbose, info=hull_info)
    elif method == "pca":
        yaw = yaw_from_pca(aligned, impl="sklearn" if lib else "closed")
    elif method == "sweep":
        if not np.isfinite(aligned[:, [0, 2]]).all():   # same contract as the PCA path
            raise ValueError(MSG_NONFINITE)
        yaw = yaw_from_sweep(aligned, yaw_steps)
    else:
        raise ValueError(f"Unknown method: {method}. Use 'pca' or 'convex_hull'")

    turned = yaw_matrix(yaw) @ aligned.T
    lo = turned.min(axis=1)
    hi = turned.max(axis=1)
    dx, dy, dz = hi - lo
    cx, cy, cz = (lo + hi) / 2
    if verbose:
        print(f"[{method}] dx={dx:.3f}, dy={dy:.3f}, dz={dz:.3f}")

    # corners are rounded to float16 in the aligned frame (:165) ...
    with np.errstate(over="ignore"):
        corners = box_corners(cx, cy, cz, dx, dy, dz, 0).astype(np.float16)
    # ... and taken back to the camera frame in float64 (:168-169)
    corners = np.dot(yaw_matrix(-yaw), corners.T).T
    corners = np.dot(corners, Rg.T)
    # centre and R_cam use Rg^T where the corners used Rg (:172-176) - kept.
    center_cam = Rg.T @ (yaw_matrix(-yaw) @ np.array([cx, cy, cz]))
    R_cam = Rg.T @ yaw_matrix(-yaw)
    return {"vertices": corners, "center_cam": center_cam, "dimension": [dz, dy, dx], "R_cam": R_cam,
            "yaw": float(yaw), "n_valid": len(aligned), "aligned": aligned, "Rg": Rg,
            "sample_idx": sample_idx, "hull_fallback": bool(hull_info.get("fallback", False))}


def estimate_bbox(in_pc, cat_name=None, ground_equ=None, method="pca", yaw_steps=None,
                  rng=None, impl="library", verbose=False, sample_idx=None):
    """Returns ``(vertices[8,3], center_cam[3], [dz,dy,dx], R_cam[3,3])``.

    ``rng`` is the legacy RandomState to draw the 500-point subsample from
    (default: the process-global ``np.random``, like the reference).
    ``impl="library"`` uses scikit-learn / SciPy like the reference;
    ``impl="closed"`` uses the closed forms the kernels implement.
    ``sample_idx`` overrides the random draw (for index-parity tests).
    """
    d = fit_details(in_pc, ground_equ, method, yaw_steps, rng, impl, verbose, sample_idx)
    return d["vertices"], d["center_cam"], d["dimension"], d["R_cam"]


# --------------------------------------------------------------------------
# a8  corner reprojection     src/util.py:227-229, combine_results.py:238-252
# --------------------------------------------------------------------------
def project_to_2d(point_3d, camera_matrix):
    h = np.dot(camera_matrix, point_3d)
    return h[:2] / h[2]


def box2d_from_corners(corners, K, W=None, H=None):
    """``(uv[8,2], bbox2D_proj[4], bbox2D_trunc[4] or None)``."""
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = np.array([project_to_2d(np.array(p), K) for p in corners])
    mnx, mny = min(p[0] for p in uv), min(p[1] for p in uv)
    mxx, mxy = max(p[0] for p in uv), max(p[1] for p in uv)
    proj = [mnx, mny, mxx, mxy]
    trunc = None if W is None else [max(0, mnx), max(0, mny), min(W, mxx), min(H, mxy)]
    return uv, proj, trunc


# --------------------------------------------------------------------------
# NumPy legacy RandomState.randint stream (src/util_3dbox.py:124 call site)
# --------------------------------------------------------------------------
class LegacyMT19937:
    """MT19937 exactly as ``np.random.RandomState(int_seed)`` drives it."""

    N, M = 624, 397

    def __init__(self, seed):
        s = int(seed) & 0xFFFFFFFF
        key = np.empty(self.N, dtype=np.uint64)
        for i in range(self.N):
            key[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
        self.key = key.astype(np.uint32)
        self.pos = self.N
        self.draws = 0

    def _regen(self):
        k = self.key.astype(np.uint64)
        N, M = self.N, self.M
        up, lo, a = np.uint64(0x80000000), np.uint64(0x7FFFFFFF), np.uint64(0x9908B0DF)

        def tw(cur, nxt, far):
            y = (cur & up) | (nxt & lo)
            return far ^ (y >> np.uint64(1)) ^ ((y & np.uint64(1)) * a)

        k[0:N - M] = tw(k[0:N - M], k[1:N - M + 1], k[M:N])
        # second stretch depends on freshly written words 227 positions back
        for start in range(N - M, N - 1, N - M):
            stop = min(start + (N - M), N - 1)
            k[start:stop] = tw(k[start:stop], k[start + 1:stop + 1], k[start - (N - M):stop - (N - M)])
        k[N - 1] = tw(k[N - 1], k[0], k[M - 1])
        self.key = k.astype(np.uint32)
        self.pos = 0

    def next_uint32(self):
        if self.pos == self.N:
            self._regen()
        y = int(self.key[self.pos])
        self.pos += 1
        self.draws += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def legacy_randint(gen, high, size):
    """``RandomState.randint(0, high, size)`` for ``1 <= high <= 2**32``:
    masked rejection on 32-bit draws, mask = smallest ``2**k-1 >= high-1``."""
    rng_max = int(high) - 1
    out = np.empty(size, dtype=np.int64)
    if rng_max == 0:
        out[:] = 0
        return out
    mask = rng_max
    for sh in (1, 2, 4, 8, 16):
        mask |= mask >> sh
    for i in range(size):
        while True:
            v = gen.next_uint32() & mask
            if v <= rng_max:
                break
        out[i] = v
    return out


# --------------------------------------------------------------------------
# Composition of SURVEY.md section 3.4: depth + masks -> packed box records
# --------------------------------------------------------------------------
FLAG_HULL_FALLBACK = 1.0   # O_PAD: the convex-hull method fell back to the PCA yaw (an addition: the reference only prints)


def pack_record(vertices, center, dims, R_cam, yaw, n_valid, status, uv, box2d, n_mask, flags=0.0):
    r = np.zeros(REC, dtype=np.float64)
    r[O_PAD] = flags
    r[O_VERT:O_VERT + 24] = np.asarray(vertices, dtype=np.float64).reshape(-1)
    r[O_CENTER:O_CENTER + 3] = center
    r[O_DIM:O_DIM + 3] = dims
    r[O_RCAM:O_RCAM + 9] = np.asarray(R_cam, dtype=np.float64).reshape(-1)
    r[O_YAW] = yaw
    r[O_NVALID] = n_valid
    r[O_STATUS] = status
    r[O_UV:O_UV + 16] = np.asarray(uv, dtype=np.float64).reshape(-1)
    r[O_BOX2D:O_BOX2D + 4] = box2d
    r[O_NMASK] = n_mask
    return r


def failed_record(status, n_valid, n_mask):
    # n_valid of a failed box is only checked where the caller knows it
    r = np.full(REC, np.nan, dtype=np.float64)
    r[O_NVALID] = n_valid
    r[O_STATUS] = status
    r[O_NMASK] = n_mask
    r[O_PAD] = 0.0
    return r


def fit_points_record(pc, K, ground=None, method="pca", yaw_steps=None, rng=None, impl="library",
                      sample_idx=None):
    """One box record from a point set (the mesh-points entry of the reference)."""
    pc = np.asarray(pc)
    n_mask = pc.shape[0]
    try:
        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            d = fit_details(pc, ground, method, yaw_steps, rng=rng, impl=impl, sample_idx=sample_idx)
    except ValueError as exc:
        return failed_record(status_of_exception(exc), np.nan, n_mask)
    uv, proj, _ = box2d_from_corners(d["vertices"], K)
    return pack_record(d["vertices"], d["center_cam"], d["dimension"], d["R_cam"], d["yaw"],
                       d["n_valid"], ST_OK, uv, proj, n_mask, FLAG_HULL_FALLBACK if d["hull_fallback"] else 0.0)


def fit_boxes(depth, K, masks, ground=None, method="pca", yaw_steps=None, seed=0, image_offset=0,
              impl="library", subsample=True):
    """``depth[B,H,W] f32, K[B,3,3], masks[B,I,H,W], ground[B,I,3]|None`` -> ``[B,I,64]`` f64.

    Per image ``b`` the legacy RandomState is re-seeded with
    ``seed + image_offset + b`` and instances draw from it in order, exactly as
    the reference would if ``np.random.seed`` were called before each image.
    ``subsample=False``: the random 500-point draw of ``src/util_3dbox.py:123-125`` is replaced by
    the identity (every masked point takes part; what ``la3d_fit_all_points`` computes).
    """
    depth = np.asarray(depth)
    masks = np.asarray(masks)
    B, I = masks.shape[:2]
    out = np.empty((B, I, REC), dtype=np.float64)
    for b in range(B):
        rng = np.random.RandomState((int(seed) + int(image_offset) + b) & 0xFFFFFFFF)
        with np.errstate(invalid="ignore", over="ignore"):
            pts = depth_to_points(depth[b][None], K[b])
        for i in range(I):
            pc = masked_points(pts, masks[b, i])
            g = None if ground is None else ground[b, i]
            every = None if subsample or pc.shape[0] <= SUBSAMPLE else np.arange(pc.shape[0])
            out[b, i] = fit_points_record(pc, K[b], g, method, yaw_steps, rng=rng, impl=impl, sample_idx=every)
    return out
