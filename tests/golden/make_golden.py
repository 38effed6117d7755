"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container (the only place ``/root/reference`` exists):

    python tests/golden/make_golden.py

It imports the reference's ``util`` / ``util_3dbox`` / ``combine_results`` in
place (``tests/live_reference.py``), feeds them deterministic inputs, stores
inputs and the reference's outputs in ``tests/golden/golden_v1.npz`` and, while
at it, checks the oracle restatement against the same outputs and prints the
worst deviation per group.  The ``.npz`` travels to the GPU box; the reference
does not.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import live_reference  # noqa: E402
from oracle import la3d_oracle as orc  # noqa: E402
from labelany3d_b200 import synth  # noqa: E402

G = {}


def put(key, value):
    G[key] = np.asarray(value)


def cloud(rng, n, center=(0.3, -0.2, 4.0), scale=(0.8, 0.3, 0.4), yaw=0.5):
    p = rng.normal(size=(n, 3)) * np.array(scale)
    c, s = np.cos(yaw), np.sin(yaw)
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return p @ rot.T + np.array(center)


def bbox_cases():
    rng = np.random.RandomState(20260101)
    cases = []
    tilt = np.array([0.08, -0.99, 0.05])
    cases.append(dict(pc=cloud(rng, 500), ground=None, seed=None))
    cases.append(dict(pc=cloud(rng, 500), ground=tilt, seed=None))
    cases.append(dict(pc=cloud(rng, 500), ground=-tilt * 3.0, seed=None))                # flipped + scaled
    cases.append(dict(pc=cloud(rng, 500), ground=np.array([0.1, -0.9, -0.2, 1.7]), seed=None))  # plane eq.
    cases.append(dict(pc=cloud(rng, 2000), ground=None, seed=7))                          # subsample live
    cases.append(dict(pc=cloud(rng, 19200, yaw=-1.1), ground=tilt, seed=11))
    cases.append(dict(pc=cloud(rng, 501), ground=tilt, seed=12))
    cases.append(dict(pc=cloud(rng, 12), ground=None, seed=None))                         # n < 20: LAPACK SVD path
    cases.append(dict(pc=cloud(rng, 19, yaw=2.0), ground=tilt, seed=None))
    cases.append(dict(pc=cloud(rng, 20, yaw=2.0), ground=tilt, seed=None))
    cases.append(dict(pc=cloud(rng, 3), ground=None, seed=None))
    cases.append(dict(pc=cloud(rng, 2), ground=None, seed=None))
    bad = cloud(rng, 300)
    bad[5, 0] = np.nan
    bad[17, 2] = np.inf
    bad[40, 1] = -np.inf
    bad[41] = np.nan
    cases.append(dict(pc=bad, ground=None, seed=None))                                    # NaN / inf rows
    cases.append(dict(pc=bad.copy(), ground=tilt, seed=None))
    cases.append(dict(pc=cloud(rng, 400).astype(np.float32), ground=tilt, seed=None))     # f32 input
    far = cloud(rng, 200, center=(100.0, -20.0, 9000.0), scale=(300.0, 5.0, 400.0))
    cases.append(dict(pc=far, ground=None, seed=None))                                    # sentinel scale
    huge = cloud(rng, 64, center=(0.0, 0.0, 9.0e4), scale=(10.0, 10.0, 10.0))
    cases.append(dict(pc=huge, ground=None, seed=None))                                   # fp16 overflow -> inf
    line = np.stack([np.linspace(-1, 1, 50), rng.normal(size=50) * 0.1, np.linspace(2, 5, 50)], 1)
    cases.append(dict(pc=line, ground=None, seed=None))                                   # collinear footprint
    cases.append(dict(pc=cloud(rng, 500, scale=(0.5, 0.5, 0.2), yaw=1.3), ground=tilt, seed=None))  # z minor
    cases.append(dict(pc=cloud(rng, 500, scale=(0.2, 0.5, 0.9), yaw=0.2), ground=tilt, seed=None))  # z major
    grid = np.stack(np.meshgrid(np.arange(5.0), [0.0, 1.0], np.arange(3.0) + 2), -1).reshape(-1, 3)
    cases.append(dict(pc=grid, ground=None, seed=None))                                   # axis-aligned lattice (ties)
    return cases


def main():
    if not live_reference.available():
        raise SystemExit("needs /root/reference")
    util, box, comb = live_reference.load()
    worst = {}

    def track(group, a, b):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        both_nan = np.isnan(a) & np.isnan(b)
        same_inf = np.isinf(a) & (a == b)
        d = np.where(both_nan | same_inf, 0.0, np.abs(a - b))
        worst[group] = max(worst.get(group, 0.0), float(np.nanmax(d)) if d.size else 0.0)
        if np.isnan(d).any():
            worst[group] = float("nan")

    # ---------------------------------------------------------------- helpers (a4)
    rng = np.random.RandomState(1)
    yaws = rng.uniform(-4, 4, 6)
    put("helpers/yaws", yaws)
    put("helpers/rotate_y", np.stack([box.rotate_y(y) for y in yaws]))
    for y, ref in zip(yaws, G["helpers/rotate_y"]):
        track("rotate_y", orc.yaw_matrix(y), ref)
    vecs = rng.normal(size=(6, 2, 3))
    vecs[0, 0] = [0, -1, 0]
    put("helpers/vec_pairs", vecs)
    put("helpers/rotation_from_vectors", np.stack([box.rotation_matrix_from_vectors(a, b) for a, b in vecs]))
    for (a, b), ref in zip(vecs, G["helpers/rotation_from_vectors"]):
        track("rotation_between", orc.rotation_between(a, b), ref)
    put("helpers/normalize", np.stack([box.normalize(a) for a, _ in vecs]))
    boxes = rng.uniform(-2, 2, (5, 7))
    boxes[:, 3:6] = np.abs(boxes[:, 3:6])
    put("helpers/box_params", boxes)
    put("helpers/box_vertices", np.stack([box.convert_box_vertices(*p) for p in boxes]))
    for p, ref in zip(boxes, G["helpers/box_vertices"]):
        track("box_corners", orc.box_corners(*p), ref)
    planes = rng.normal(size=(4, 7))
    put("helpers/plane_args", planes)
    put("helpers/plane_dist", [box.point_to_plane_distance(p[:4], p[4], p[5], p[6]) for p in planes])
    for p, ref in zip(planes, G["helpers/plane_dist"]):
        track("plane_dist", orc.point_to_plane_distance(p[:4], p[4], p[5], p[6]), ref)

    # ---------------------------------------------------------------- depth lift (a1)
    H, W = 24, 32
    Kp = np.array([[0.9 * W, 0, W / 2], [0, 0.9 * W, H / 2], [0, 0, 1.0]])
    Ks = np.array([[31.7, 0.3, 15.1], [0.02, 29.9, 12.6], [1e-4, -2e-4, 1.0]])
    Rr = box.rotation_matrix_from_vectors(np.array([0.0, 0, 1]), np.array([0.2, -0.1, 0.9]))
    tt = np.array([0.5, -1.5, 2.0])
    d0 = rng.uniform(0.5, 8, (1, H, W)).astype(np.float32)
    d1 = d0.copy()
    d1[0, 0, 0] = np.inf
    d1[0, 3, 4] = np.nan
    d1[0, 5, 5] = 0.0
    d1[0, 7, 1] = -np.inf
    d1[0, 9, 9] = 10000.0
    d2 = rng.uniform(0.5, 8, (3, H, W)).astype(np.float32)  # bs > 1: only element 0 comes back
    lift = [(d0, Kp, None, None), (d0, Ks, Rr, tt), (d1, Kp, None, None), (d1, Ks, Rr, tt), (d2, Kp, None, tt),
            (d0, Kp, Rr, None)]
    put("lift/n", len(lift))
    for i, (d, K, R, t) in enumerate(lift):
        with np.errstate(invalid="ignore"):
            ref = util.depth_to_points(d, K, R, t)
            mine = orc.depth_to_points(d, K, R, t)
            closed = orc.depth_to_points_closed(d[0], np.linalg.inv(K), R, t)
        put(f"lift/{i}/depth", d)
        put(f"lift/{i}/K", K)
        put(f"lift/{i}/Kinv", np.linalg.inv(K))
        put(f"lift/{i}/R", np.zeros(0) if R is None else R)
        put(f"lift/{i}/t", np.zeros(0) if t is None else t)
        put(f"lift/{i}/out", ref)
        assert ref.dtype == np.float64 and ref.shape == (H, W, 3)
        track("lift(oracle)", mine, ref)
        track("lift(closed form)", closed, ref)

    # ---------------------------------------------------------------- box from points (a3, a5, a6)
    cases = bbox_cases()
    put("bbox/n", len(cases))
    for i, c in enumerate(cases):
        pc, g, seed = c["pc"], c["ground"], c["seed"]
        put(f"bbox/{i}/pc", pc)
        put(f"bbox/{i}/ground", np.zeros(0) if g is None else g)
        put(f"bbox/{i}/seed", -1 if seed is None else seed)
        if seed is not None:
            put(f"bbox/{i}/sample_idx", np.random.RandomState(seed).randint(0, pc.shape[0], 500))
        for method in ("pca", "convex_hull"):
            if seed is not None:
                np.random.seed(seed)
            try:
                with live_reference.quiet(), np.errstate(invalid="ignore", over="ignore"):
                    v, ctr, dim, Rc = box.estimate_bbox(pc, "thing", g, method)
                status = orc.ST_OK
            except ValueError as exc:
                status = orc.status_of_exception(exc)
                v, ctr, dim, Rc = np.full((8, 3), np.nan), np.full(3, np.nan), [np.nan] * 3, np.full((3, 3), np.nan)
            put(f"bbox/{i}/{method}/status", status)
            put(f"bbox/{i}/{method}/vertices", v)
            put(f"bbox/{i}/{method}/center", ctr)
            put(f"bbox/{i}/{method}/dims", np.array(dim, dtype=np.float64))
            put(f"bbox/{i}/{method}/R_cam", Rc)
            for impl in ("library", "closed"):
                rs = None if seed is None else np.random.RandomState(seed)
                try:
                    with np.errstate(invalid="ignore", over="ignore"):
                        ov, oc, od, oR = orc.estimate_bbox(pc, None, g, method, rng=rs, impl=impl)
                    assert status == orc.ST_OK, (i, method, impl, status)
                except ValueError as exc:
                    assert orc.status_of_exception(exc) == status, (i, method, impl, exc)
                    continue
                for name, a, b in (("vertices", ov, v), ("center", oc, ctr), ("dims", od, dim), ("R_cam", oR, Rc)):
                    track(f"bbox[{method},{impl}] {name}", a, b)
    # error behaviour
    for name, pc in (("one_point", np.array([[0.1, 0.2, 3.0]])), ("all_nan", np.full((4, 3), np.nan))):
        try:
            with live_reference.quiet():
                box.estimate_bbox(pc, None, None, "pca")
            msg = ""
        except Exception as exc:  # noqa: BLE001
            msg = f"{type(exc).__name__}: {exc}"
        put(f"errors/{name}", np.array(msg))
    try:
        box.estimate_bbox(np.zeros((5, 3)), None, None, "nope")
    except Exception as exc:  # noqa: BLE001
        put("errors/bad_method", np.array(f"{type(exc).__name__}: {exc}"))
    try:
        with live_reference.quiet(), np.errstate(invalid="ignore"):
            box.estimate_bbox(cloud(rng, 50), None, np.array([0.0, -2.0, 0.0]), "pca")
        put("errors/parallel_ground", np.array(""))
    except Exception as exc:  # noqa: BLE001
        put("errors/parallel_ground", np.array(f"{type(exc).__name__}: {exc}"))

    # ---------------------------------------------------------------- projection (a8)
    pts = rng.normal(size=(8, 3)) + np.array([0, 0, 4.0])
    Kq = np.array([[576.0, 0, 320], [0, 576.0, 240], [0, 0, 1]])
    put("proj/pts", pts)
    put("proj/K", Kq)
    put("proj/uv_util", np.stack([util.project_to_2d(p, Kq) for p in pts]))
    put("proj/uv_combine", np.stack([comb.project_to_2d(p, Kq) for p in pts]))
    uv, proj, trunc = orc.box2d_from_corners(pts, Kq, 640, 480)
    track("project_to_2d", uv, G["proj/uv_util"])
    track("project_to_2d", uv, G["proj/uv_combine"])
    ref_uv = G["proj/uv_combine"]
    ref_proj = [ref_uv[:, 0].min(), ref_uv[:, 1].min(), ref_uv[:, 0].max(), ref_uv[:, 1].max()]
    put("proj/bbox2D_proj", ref_proj)
    put("proj/bbox2D_trunc", [max(0, ref_proj[0]), max(0, ref_proj[1]), min(640, ref_proj[2]), min(480, ref_proj[3])])
    track("bbox2D", proj, ref_proj)
    track("bbox2D", trunc, G["proj/bbox2D_trunc"])

    # ---------------------------------------------------------------- legacy randint stream
    rcases = [(0, 501), (1, 19200), (1234, 37000), (2 ** 32 - 1, 512), (77, 513), (5, 147456), (99, 1)]
    put("rng/cases", np.array(rcases, dtype=np.int64))
    for i, (seed, high) in enumerate(rcases):
        rs = np.random.RandomState(seed)
        a = rs.randint(0, high, 500)
        b = rs.randint(0, high + 3, 500)      # second call continues the same stream
        put(f"rng/{i}/first", a)
        put(f"rng/{i}/second", b)
        gen = orc.LegacyMT19937(seed)
        track("legacy_randint", orc.legacy_randint(gen, high, 500), a)
        track("legacy_randint", orc.legacy_randint(gen, high + 3, 500), b)

    # ---------------------------------------------------------------- composed path (section 3.4)
    import torch  # noqa: F401  (synth uses torch)
    B, H, W, I = 3, 96, 128, 4
    depth, K, masks, ground = synth.make_inputs(B, H, W, I, seed=4321, device="cpu", area=(0.05, 0.2))
    depth, K, masks, ground = depth.numpy(), K.numpy(), masks.numpy(), ground.numpy()
    masks[0, 3] = False
    masks[0, 3, 40:50, 60:80] = True          # 200 px: no subsample
    masks[1, 2] = False                       # empty mask -> ValueError in the reference
    masks[2, 1] = False
    masks[2, 1, 10, 10] = True                # single pixel -> scikit-learn refuses
    depth[1, 20:24, 30:60] = np.inf           # invalid depth inside masks
    depth[2, 50, 64:90] = np.nan
    seed = 1234
    put("scene/depth", depth)
    put("scene/K", K)
    put("scene/masks", np.packbits(masks, axis=-1))
    put("scene/shape", [B, I, H, W])
    put("scene/ground", ground)
    put("scene/seed", seed)
    for use_ground in (0, 1):
        for method in ("pca", "convex_hull"):
            rec = np.full((B, I, orc.REC), np.nan)
            ranks = np.full((B, I, 500), -1, dtype=np.int64)
            for b in range(B):
                np.random.seed(seed + b)
                shadow = np.random.RandomState(seed + b)   # same stream, to record the draws
                with np.errstate(invalid="ignore"):
                    pts3 = util.depth_to_points(depth[b][None], K[b])
                for i in range(I):
                    pc = pts3[masks[b, i]]
                    n_mask = pc.shape[0]
                    if n_mask > 500:
                        ranks[b, i] = shadow.randint(0, n_mask, 500)
                    g = ground[b, i] if use_ground else None
                    try:
                        with live_reference.quiet(), np.errstate(invalid="ignore", over="ignore"):
                            v, ctr, dim, Rc = box.estimate_bbox(pc, None, g, method)
                    except ValueError as exc:
                        rec[b, i] = orc.failed_record(orc.status_of_exception(exc), np.nan, n_mask)
                        continue
                    uv = np.stack([comb.project_to_2d(np.array(p), K[b]) for p in v])
                    proj = [uv[:, 0].min(), uv[:, 1].min(), uv[:, 0].max(), uv[:, 1].max()]
                    r = orc.pack_record(v, ctr, dim, Rc, np.nan, np.nan, orc.ST_OK, uv, proj, n_mask)
                    rec[b, i] = r
            put(f"scene/g{use_ground}/{method}/records", rec)
            put(f"scene/g{use_ground}/{method}/ranks", ranks)
            for impl in ("library", "closed"):
                mine = orc.fit_boxes(depth, K, masks, ground if use_ground else None, method, seed=seed, impl=impl)
                sel = np.ones(orc.REC, dtype=bool)
                sel[[orc.O_YAW, orc.O_NVALID]] = False    # not observable through the reference API
                track(f"scene[{method},{impl}]", mine[..., sel], rec[..., sel])

    # ---------------------------------------------------------------- cam_utils helpers
    import importlib.util
    spec = importlib.util.spec_from_file_location("_la3d_ref_cam", os.path.join(live_reference.REF_SRC, "cam_utils.py"))
    cam = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cam)
    orbit_args = [(10.0, 20.0, 1.0, 1, 1), (-30.0, 140.0, 2.5, 1, 0), (0.3, -1.2, 1.0, 0, 1), (0.0, 0.0, 1.0, 1, 0)]
    put("cam/orbit_args", np.array(orbit_args))
    put("cam/orbit", np.stack([cam.orbit_camera(e, a, r, bool(d), None, bool(o)) for e, a, r, d, o in orbit_args]))
    tgt = np.array([0.5, -0.25, 1.0], dtype=np.float32)
    put("cam/orbit_target", np.stack([cam.orbit_camera(e, a, r, bool(d), tgt, bool(o)) for e, a, r, d, o in orbit_args]))
    vv = rng.normal(size=(5, 3))
    put("cam/vecs", vv)
    put("cam/length", cam.length(vv))
    put("cam/safe_normalize", cam.safe_normalize(vv))

    # ---------------------------------------------------------------- scene driver (a9) + draw_cube (a10)
    # trimesh is absent from this image: the reference's driver runs unmodified against a
    # stand-in that serves pre-sampled points, so its directory / JSON logic is pinned too.
    import json
    import tempfile
    import types
    import cv2
    scene = tempfile.mkdtemp()
    os.makedirs(os.path.join(scene, "reconstruction"))
    objs = [("0", "chair", 500), ("1", "dining_table", 500), ("2", "bad", 0), ("3", "cup", 500)]
    srng = np.random.RandomState(99)
    clouds = {}
    for oid, cat, n in objs:
        name = f"{oid}_{cat}"
        open(os.path.join(scene, "reconstruction", name + ".glb"), "w").close()
        clouds[name] = cloud(srng, n, center=(srng.uniform(-1, 1), 0.2, srng.uniform(3, 6)), yaw=srng.uniform(-2, 2)) if n else None
        up = np.array([srng.normal() * 0.1, -1.0 + srng.normal() * 0.05, srng.normal() * 0.1, 0.0]) * 1.7
        np.save(os.path.join(scene, "reconstruction", name + "_canonical_upright.npy"), up)
        put(f"driver/{name}/points", np.zeros((0, 3)) if clouds[name] is None else clouds[name])
        put(f"driver/{name}/upright", up)
    open(os.path.join(scene, "reconstruction", "full_scene.glb"), "w").close()
    put("driver/names", np.array([f"{o}_{c}" for o, c, _ in objs]))

    class FakeMesh:
        def __init__(self, pts):
            self.pts = pts
            self.is_empty = pts is None
            self.area = 0.0 if pts is None else 1.0
            self.faces = [] if pts is None else [0]

        def sample(self, n):
            assert n == 500
            return self.pts

    fake = types.ModuleType("trimesh")
    fake.Scene = type("Scene", (), {})
    fake.load = lambda path: FakeMesh(clouds[os.path.basename(path)[:-4]])
    fake.points = types.SimpleNamespace(PointCloud=lambda p: types.SimpleNamespace(vertices=p))
    box.trimesh = fake
    for method in ("pca", "convex_hull"):
        with live_reference.quiet():
            lst = box.save_3d_with_ground_alignment_bbox(scene, method)
        with open(os.path.join(scene, "3dbbox_ground.json")) as f:
            assert json.load(f) == lst
        put(f"driver/json_{method}", np.array(json.dumps(lst)))
    # draw_cube on the pca boxes
    Wd, Hd = 160, 120
    Kd = np.array([[0.9 * Wd, 0, Wd / 2], [0, 0.9 * Wd, Hd / 2], [0, 0, 1.0]])
    with open(os.path.join(scene, "cam_params.json"), "w") as f:
        json.dump({"K": Kd.tolist(), "H": Hd, "W": Wd}, f)
    img = (np.random.RandomState(5).rand(Hd, Wd, 3) * 255).astype(np.uint8)
    cv2.imwrite(os.path.join(scene, "input.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    with live_reference.quiet():
        box.save_3d_with_ground_alignment_bbox(scene, "pca")
    util.draw_cube(scene, is_ground=True)
    put("driver/K", Kd)
    put("driver/input_rgb", img)
    put("driver/vis_bgr", cv2.imread(os.path.join(scene, "vis_3dbox.png")))

    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **G)
    print(f"wrote {out}: {len(G)} arrays, {os.path.getsize(out) / 1024:.1f} KiB")
    for k in sorted(worst):
        print(f"  oracle vs reference  {k:40s} max|diff| = {worst[k]:.3e}")


if __name__ == "__main__":
    main()
