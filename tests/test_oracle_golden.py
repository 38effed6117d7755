"""The oracle restatement replayed against the reference's own outputs (golden vectors).

These are the pins that make ``oracle/la3d_oracle.py`` trustworthy on the GPU box,
where the reference itself is not available.  ``impl="library"`` (scikit-learn /
SciPy, like the reference) must be bit-exact; ``impl="closed"`` (the closed forms
the CUDA kernels implement) must agree to rounding.
"""

import numpy as np
import pytest

from conftest import close, opt
from oracle import la3d_oracle as orc


def test_helpers(golden):
    for y, ref in zip(golden["helpers/yaws"], golden["helpers/rotate_y"]):
        np.testing.assert_array_equal(orc.yaw_matrix(y), ref)
    for (a, b), ref, nrm in zip(golden["helpers/vec_pairs"], golden["helpers/rotation_from_vectors"],
                                golden["helpers/normalize"]):
        np.testing.assert_array_equal(orc.rotation_between(a, b), ref)
        np.testing.assert_array_equal(orc.unit(a), nrm)
    for p, ref in zip(golden["helpers/box_params"], golden["helpers/box_vertices"]):
        np.testing.assert_array_equal(orc.box_corners(*p), ref)
    for p, ref in zip(golden["helpers/plane_args"], golden["helpers/plane_dist"]):
        assert orc.point_to_plane_distance(p[:4], p[4], p[5], p[6]) == ref


def test_depth_lift(golden):
    for i in range(int(golden["lift/n"])):
        d, K = golden[f"lift/{i}/depth"], golden[f"lift/{i}/K"]
        R, t = opt(golden[f"lift/{i}/R"]), opt(golden[f"lift/{i}/t"])
        ref = golden[f"lift/{i}/out"]
        with np.errstate(invalid="ignore"):
            out = orc.depth_to_points(d, K, R, t)
            closed = orc.depth_to_points_closed(d[0], golden[f"lift/{i}/Kinv"], R, t)
        assert out.dtype == np.float64
        np.testing.assert_array_equal(out, ref)
        if R is None:
            np.testing.assert_array_equal(closed, ref)       # K-only calls: the closed form is bit-exact
        else:
            close(closed, ref, 4e-16 * max(1.0, np.nanmax(np.abs(ref[np.isfinite(ref)]))) * 4)


@pytest.mark.parametrize("method", ["pca", "convex_hull"])
@pytest.mark.parametrize("impl", ["library", "closed"])
def test_bbox_from_points(golden, method, impl):
    for i in range(int(golden["bbox/n"])):
        pc, g = golden[f"bbox/{i}/pc"], opt(golden[f"bbox/{i}/ground"])
        seed = int(golden[f"bbox/{i}/seed"])
        status = int(golden[f"bbox/{i}/{method}/status"])
        rs = None if seed < 0 else np.random.RandomState(seed)
        try:
            with np.errstate(invalid="ignore", over="ignore"):
                v, c, d, R = orc.estimate_bbox(pc, "thing", g, method, rng=rs, impl=impl)
            got = orc.ST_OK
        except ValueError as exc:
            got = orc.status_of_exception(exc)
        assert got == status, (i, got, status)
        if status != orc.ST_OK:
            continue
        ref = [golden[f"bbox/{i}/{method}/{k}"] for k in ("vertices", "center", "dims", "R_cam")]
        if impl == "library":
            for a, b in zip((v, c, d, R), ref):
                np.testing.assert_array_equal(np.asarray(a, dtype=np.float64), b)
        else:
            fin = np.abs(pc[np.isfinite(pc)])
            # X^T X - n mu mu^T cancels ~(|mu|/spread)^2 digits in sklearn and here alike,
            # so the allowance grows with the cloud's distance; 1e-4 abs is the product bar
            tol = max(1e-11, 2e-12 * fin.max())
            for a, b in zip((v, c, d, R), ref):
                close(a, b, tol)


def test_bbox_errors(golden):
    for name, pc, ground, method in (
            ("one_point", np.array([[0.1, 0.2, 3.0]]), None, "pca"),
            ("all_nan", np.full((4, 3), np.nan), None, "pca"),
            ("bad_method", np.zeros((5, 3)), None, "nope")):
        expect = str(golden[f"errors/{name}"])
        for impl in ("library", "closed"):
            with pytest.raises(ValueError) as info:
                orc.estimate_bbox(pc, None, ground, method, impl=impl)
            kind, msg = expect.split(": ", 1)
            assert kind == "ValueError"
            if impl == "library" or name != "one_point":
                assert str(info.value) == msg
            else:
                assert str(info.value).startswith("n_components=2 must be between 0 and min(n_samples, n_features)=1")
    # ground exactly parallel to (0,-1,0): 0/0 in Rodrigues -> all NaN -> "No valid points"
    assert str(golden["errors/parallel_ground"]) == "ValueError: " + orc.MSG_NO_VALID
    with pytest.raises(ValueError, match="No valid points"), np.errstate(invalid="ignore"):
        orc.estimate_bbox(np.random.RandomState(0).normal(size=(50, 3)), None, np.array([0.0, -2.0, 0.0]))


def test_projection(golden):
    uv, proj, trunc = orc.box2d_from_corners(golden["proj/pts"], golden["proj/K"], 640, 480)
    np.testing.assert_array_equal(uv, golden["proj/uv_util"])
    np.testing.assert_array_equal(uv, golden["proj/uv_combine"])
    np.testing.assert_array_equal(proj, golden["proj/bbox2D_proj"])
    np.testing.assert_array_equal(trunc, golden["proj/bbox2D_trunc"])


def test_legacy_randint_stream(golden):
    for i, (seed, high) in enumerate(golden["rng/cases"]):
        gen = orc.LegacyMT19937(int(seed))
        np.testing.assert_array_equal(orc.legacy_randint(gen, int(high), 500), golden[f"rng/{i}/first"])
        np.testing.assert_array_equal(orc.legacy_randint(gen, int(high) + 3, 500), golden[f"rng/{i}/second"])
        # and NumPy itself still produces the stream the goldens were minted with
        rs = np.random.RandomState(int(seed))
        np.testing.assert_array_equal(rs.randint(0, int(high), 500), golden[f"rng/{i}/first"])


def scene_inputs(golden):
    B, I, H, W = (int(x) for x in golden["scene/shape"])
    masks = np.unpackbits(golden["scene/masks"], axis=-1)[..., :W].astype(bool)
    return golden["scene/depth"], golden["scene/K"], masks, golden["scene/ground"], int(golden["scene/seed"])


@pytest.mark.parametrize("method", ["pca", "convex_hull"])
@pytest.mark.parametrize("use_ground", [0, 1])
@pytest.mark.parametrize("impl", ["library", "closed"])
def test_composed_scene(golden, method, use_ground, impl):
    depth, K, masks, ground, seed = scene_inputs(golden)
    ref = golden[f"scene/g{use_ground}/{method}/records"]
    rec = orc.fit_boxes(depth, K, masks, ground if use_ground else None, method, seed=seed, impl=impl)
    sel = np.ones(orc.REC, dtype=bool)
    sel[[orc.O_YAW, orc.O_NVALID]] = False       # not observable through the reference API
    if impl == "library":
        np.testing.assert_array_equal(rec[..., sel], ref[..., sel])
    else:
        close(rec[..., sel], ref[..., sel], 1e-9)
    # integer work: mask counts, statuses and the sampled ranks are exact
    np.testing.assert_array_equal(rec[..., orc.O_NMASK], orc.mask_counts(masks))
    np.testing.assert_array_equal(rec[..., orc.O_STATUS], ref[..., orc.O_STATUS])
    assert set(np.unique(ref[..., orc.O_STATUS])) == {orc.ST_OK, orc.ST_NO_VALID, orc.ST_PCA_UNDEFINED}
    ranks = golden[f"scene/g{use_ground}/{method}/ranks"]
    for b in range(masks.shape[0]):
        gen = orc.LegacyMT19937(seed + b)
        for i in range(masks.shape[1]):
            n = int(masks[b, i].sum())
            if n > orc.SUBSAMPLE:
                np.testing.assert_array_equal(orc.legacy_randint(gen, n, 500), ranks[b, i])


def test_sweep_is_self_consistent():
    """The sweep has no reference; check the restatement against brute force properties."""
    rng = np.random.RandomState(3)
    pc = rng.normal(size=(300, 3)) * np.array([1.0, 0.2, 0.3])
    ang = 0.4
    rot = orc.yaw_matrix(ang)
    pc = pc @ rot.T
    for steps in (36, 360):
        yaw = orc.yaw_from_sweep(pc, steps)
        k = yaw / ((np.pi / 2) / steps)
        assert abs(k - round(k)) < 1e-9 and 0 <= round(k) < steps
        areas = []
        for j in range(steps):
            r = orc.yaw_matrix(j * (np.pi / 2) / steps) @ pc.T
            areas.append((r[0].max() - r[0].min()) * (r[2].max() - r[2].min()))
        assert int(round(k)) == int(np.argmin(areas))
        v, c, d, R = orc.estimate_bbox(pc, None, None, "sweep", steps)
        assert d[2] * d[0] == pytest.approx(min(areas), rel=1e-12)
