"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the host helpers reproduce the reference's values, and the product never imports
the oracle.  No compute call is made on the GPU here."""

import ctypes
import os
import re
import sys

import numpy as np
import pytest

import labelany3d_b200
from labelany3d_b200 import _lib, build, records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dropin():
    path = labelany3d_b200.dropin_path()
    if path not in sys.path:
        sys.path.insert(0, path)
    for name in ("util", "util_3dbox", "cam_utils"):
        sys.modules.pop(name, None)
    import cam_utils
    import util
    import util_3dbox
    assert util_3dbox.__file__.startswith(path)
    return util, util_3dbox, cam_utils


def test_library_exports_every_header_symbol():
    path = build.build()
    assert os.path.isfile(path)
    header = open(os.path.join(ROOT, "include", "la3d.h")).read()
    declared = set(re.findall(r"\b(la3d_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().la3d_version() == 200
    # layout helpers are pure host functions
    assert _lib.load().la3d_chunks_per_plane(480, 640) == 600
    assert _lib.load().la3d_words_per_plane(480, 640) == 9600
    assert _lib.load().la3d_chunks_per_plane(7, 13) == 1
    assert _lib.load().la3d_fit_workspace_bytes(256, 8, 480, 640) > 256 * 8 * 9600 * 4


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/la3d.h compiles as strict C99 (no C++ or torch types in the boundary) and a C program links
    against the shared library and gets versions, sizes and argument errors through it (examples/c_host.c)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "labelany3d_b200", "lib")
    exe = str(tmp_path / "c_host")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "examples", "c_host.c"), "-L", lib_dir, "-lla3d_sm100a", "-Wl,-rpath," + lib_dir, "-o", exe]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "la3d version" in run.stdout and "null pointers -> -1 (la3d_fit_boxes: null pointer)" in run.stdout


def test_c_abi_rejects_bad_arguments_before_touching_the_gpu():
    """Argument checks come first in every entry point: null pointers / bad shapes return LA3D_EINVAL
    with a message naming the function, without a CUDA call (so this runs on the CPU box)."""
    lib = _lib.load()
    EINVAL = -1
    calls = {
        "la3d_depth_lift": lambda: lib.la3d_depth_lift(None, None, 9, 0, None, None, 1, 4, 4, None, 0, None),
        "la3d_mask_scan": lambda: lib.la3d_mask_scan(None, 1, 4, 4, 1, None, None, None),
        "la3d_mask_scan_thin": lambda: lib.la3d_mask_scan_thin(None, 1, 16, 32, 1, None, None, 2, 2, None),
        "la3d_mask_stats": lambda: lib.la3d_mask_stats(None, 1, 4, 4, None, None, None),
        "la3d_mask_overlap": lambda: lib.la3d_mask_overlap(None, None, 1, 1, 4, 4, None, None),
        "la3d_fit_prepare": lambda: lib.la3d_fit_prepare(None, None, 1, 1, 0, 0, None, 0, None),
        "la3d_sample_ranks": lambda: lib.la3d_sample_ranks(None, None, 1, 1, 4, 4, None, None, None),
        "la3d_fit_scanned": lambda: lib.la3d_fit_scanned(None, None, None, None, None, 1, 1, 4, 4, 0, 0, None, 0, None),
        "la3d_fit_boxes": lambda: lib.la3d_fit_boxes(None, None, None, None, 1, 1, 4, 4, 1, 0, 0, 0, 0, None, 0, None, 0, None),
        "la3d_fit_boxes_to": lambda: lib.la3d_fit_boxes_to(None, None, None, None, 1, 1, 4, 4, 1, 0, 0, 0, 0, None, 0, None, None),
        "la3d_fit_scanned_to": lambda: lib.la3d_fit_scanned_to(None, None, None, None, None, 1, 1, 4, 4, 0, 0, None, None),
        "la3d_fit_boxes_rle_to": lambda: lib.la3d_fit_boxes_rle_to(None, None, None, 0, None, None, None, 1, 1, 4, 4, 0, 0, 0, 0, None, 0, None, None, None),
        "la3d_fit_all_points_to": lambda: lib.la3d_fit_all_points_to(None, None, None, 1, 1, 4, 4, 0, 0, None, None),
        "la3d_fit_boxes_all_to": lambda: lib.la3d_fit_boxes_all_to(None, None, None, None, 1, 1, 4, 4, 1, 0, 0, None, 0, None, None),
        "la3d_peer_barrier": lambda: lib.la3d_peer_barrier(None, 0, 1, 1, None, None),
        "la3d_peer_signal": lambda: lib.la3d_peer_signal(None, 0, 1, 1, None),
        "la3d_peer_wait": lambda: lib.la3d_peer_wait(None, 0, 1, 1, None, None),
        "la3d_fit_points": lambda: lib.la3d_fit_points(None, None, None, None, None, 1, 0, 0, None, 0, None),
        "la3d_project_points": lambda: lib.la3d_project_points(None, None, None, 1, None, None),
        "la3d_iou_matrix": lambda: lib.la3d_iou_matrix(None, None, None, None, None, 1, None, None),
        "la3d_box2d_from_corners": lambda: lib.la3d_box2d_from_corners(None, None, None, None, 1, None, None, None),
        "la3d_rle_decode": lambda: lib.la3d_rle_decode(None, None, 1, 4, 4, 0, None, None, None, None, None),
        "la3d_fit_boxes_bits": lambda: lib.la3d_fit_boxes_bits(None, None, None, None, None, 1, 1, 4, 4, 0, 0, 0, 0, None, 0, None, 0, None),
        "la3d_fit_boxes_rle": lambda: lib.la3d_fit_boxes_rle(None, None, None, 0, None, None, None, 1, 1, 4, 4, 0, 0, 0, 0, None, 0, None, None, 0, None),
        "la3d_fit_all_points": lambda: lib.la3d_fit_all_points(None, None, None, 1, 1, 4, 4, None, 0, None),
        "la3d_fit_boxes_all": lambda: lib.la3d_fit_boxes_all(None, None, None, None, 1, 1, 4, 4, 1, None, 0, None, 0, None),
        "la3d_ransac_subset_fit": lambda: lib.la3d_ransac_subset_fit(None, None, None, 1, None, None),
        "la3d_ransac_classify": lambda: lib.la3d_ransac_classify(None, None, 1, 1.0, 1.0, None, None),
        "la3d_scale_fill": lambda: lib.la3d_scale_fill(None, None, 1, 1.0, 1.0, None, None),
        "la3d_masked_ratio_median": lambda: lib.la3d_masked_ratio_median(None, None, None, None, 1, 1, 4, 4, None, None, None),
    }
    for name, call in calls.items():
        assert call() == EINVAL, name
        assert b"null pointer" in lib.la3d_last_error(), (name, lib.la3d_last_error())
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert lib.la3d_mask_scan(p, 0, 4, 4, 1, p, p, None) == EINVAL and b"non-positive shape" in lib.la3d_last_error()
    assert lib.la3d_mask_scan_thin(p, 1, 3, 5, 1, p, p, 2, 2, None) == EINVAL and b"multiple of 512" in lib.la3d_last_error()
    assert lib.la3d_depth_lift(p, p, 5, 0, None, None, 1, 4, 4, p, 0, None) == EINVAL and b"k_stride" in lib.la3d_last_error()
    assert lib.la3d_fit_scanned(p, p, p, p, p, 1, 1, 4, 4, 2, 0, p, 0, None) == EINVAL and b"yaw_steps" in lib.la3d_last_error()
    assert lib.la3d_peer_barrier(ctypes.cast(p, ctypes.c_void_p), 3, 2, 1, None, None) == EINVAL
    # sinks: destinations, alignment, the synchronisation block
    from labelany3d_b200 import _lib as L
    bad = [L.make_sink([], 0), L.make_sink([p + 4], 0), L.make_sink([p], 0, flags=[p], counter=None, status=None, epoch=1),
           L.make_sink([p], 0, flags=[p], counter=p, status=None, epoch=0), L.make_sink([p, p], 0, flags=[p, p], counter=p, epoch=1, rank=2)]
    for sink in bad:
        assert lib.la3d_fit_boxes_to(p, p, p, None, 1, 1, 4, 4, 1, 0, 0, 0, 0, p, 1 << 20, ctypes.byref(sink), None) == EINVAL
    assert lib.la3d_prep_bytes(0, 4) == 0 and lib.la3d_fit_workspace_bytes(1, 0, 4, 4) == 0
    assert lib.la3d_prep_bytes(256, 8) > 256 * 8 * 1024 * 4


def test_record_layout_matches_header_and_oracle():
    from oracle import la3d_oracle as orc
    header = open(os.path.join(ROOT, "include", "la3d.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define LA3D_(O_[A-Z0-9]+|REC|ST_[A-Z_]+|SUBSAMPLE|METHOD_[A-Z_]+)\s+(-?\d+)", header)}
    for name in ("O_VERT", "O_CENTER", "O_DIM", "O_RCAM", "O_YAW", "O_NVALID", "O_STATUS", "O_UV", "O_BOX2D", "O_NMASK",
                 "O_PAD", "REC", "ST_OK", "ST_NO_VALID", "ST_PCA_UNDEFINED", "ST_BAD_METHOD", "ST_NONFINITE", "SUBSAMPLE"):
        assert defs[name] == getattr(records, name) == getattr(orc, name), name
    assert records.METHODS == {"pca": defs["METHOD_PCA"], "convex_hull": defs["METHOD_CONVEX_HULL"],
                               "sweep": defs["METHOD_SWEEP"]}


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "labelany3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(dirpath, f)
                assert "la3d_oracle" not in text, os.path.join(dirpath, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.La3dError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from labelany3d_b200 import ops
    with pytest.raises(TypeError, match="CUDA tensor"):
        ops.depth_lift(torch.zeros(1, 4, 4), torch.eye(3, dtype=torch.float64))
    with pytest.raises(TypeError, match="CUDA tensor"):
        ops.mask_scan(torch.zeros(1, 1, 4, 4, dtype=torch.bool))


def test_status_errors_carry_the_reference_messages(golden):
    assert str(records.status_error(records.ST_NO_VALID)) == str(golden["errors/all_nan"]).split(": ", 1)[1]
    assert str(records.status_error(records.ST_BAD_METHOD, "nope")) == str(golden["errors/bad_method"]).split(": ", 1)[1]
    assert str(records.status_error(records.ST_PCA_UNDEFINED, n_valid=1)) == str(golden["errors/one_point"]).split(": ", 1)[1]
    assert records.status_error(records.ST_OK) is None
    assert records.bbox2d_trunc([-5.0, 3.0, 700.0, 500.0], 640, 480) == [0, 3.0, 640, 480]


def test_geometry_helpers(dropin, golden):
    _, box, _ = dropin
    for y, ref in zip(golden["helpers/yaws"], golden["helpers/rotate_y"]):
        np.testing.assert_array_equal(box.rotate_y(y), ref)
    for (a, b), ref, nrm in zip(golden["helpers/vec_pairs"], golden["helpers/rotation_from_vectors"],
                                golden["helpers/normalize"]):
        np.testing.assert_array_equal(box.rotation_matrix_from_vectors(a, b), ref)
        np.testing.assert_array_equal(box.normalize(a), nrm)
    assert box.normalize(np.zeros(3)).tolist() == [0, 0, 0]
    for p, ref in zip(golden["helpers/box_params"], golden["helpers/box_vertices"]):
        np.testing.assert_array_equal(box.convert_box_vertices(*p), ref)
    for p, ref in zip(golden["helpers/plane_args"], golden["helpers/plane_dist"]):
        assert box.point_to_plane_distance(p[:4], p[4], p[5], p[6]) == ref


def test_cam_utils(dropin, golden):
    _, _, cam = dropin
    tgt = np.array([0.5, -0.25, 1.0], dtype=np.float32)
    for (e, a, r, d, o), ref, ref_t in zip(golden["cam/orbit_args"], golden["cam/orbit"], golden["cam/orbit_target"]):
        np.testing.assert_array_equal(cam.orbit_camera(e, a, r, bool(d), None, bool(o)), ref)
        np.testing.assert_array_equal(cam.orbit_camera(e, a, r, bool(d), tgt, bool(o)), ref_t)
    np.testing.assert_array_equal(cam.length(golden["cam/vecs"]), golden["cam/length"])
    np.testing.assert_array_equal(cam.safe_normalize(golden["cam/vecs"]), golden["cam/safe_normalize"])
    import torch
    v = torch.as_tensor(golden["cam/vecs"])
    np.testing.assert_allclose(cam.length(v).numpy(), golden["cam/length"], rtol=1e-15)   # reference's torch branch is broken
    assert cam.depth_to_points is dropin[0].depth_to_points


def test_signatures_match_the_reference(dropin):
    import inspect
    util, box, cam = dropin
    want = {
        (util, "depth_to_points"): "(depth, K=None, R=None, t=None)",
        (util, "project_to_2d"): "(point_3d, camera_matrix)",
        (util, "draw_cube"): "(scene_dir, is_ground=False)",
        (util, "analyze_mask"): "(mask, image_size, scale_threshold=100, boundary_threshold=10)",
        (util, "get_maximum_height"): "(binary_mask)",
        (util, "read_bounding_boxes_segmentations"): "(annotations_path_or_list, image_size)",
        (util, "create_boolean_mask_from_polygon"): "(image_shape, segmentation)",
        (util, "replace_categories_with_supercategories"): "(category_ids, json_file_path=None)",
        (util, "align_to_depth_match"): "(mask, depth_map, object_name, project_root, model)",
        (box, "normalize"): "(v)",
        (box, "rotate_y"): "(yaw)",
        (box, "rotation_matrix_from_vectors"): "(vec1, vec2)",
        (box, "point_to_plane_distance"): "(plane, x, y, z)",
        (box, "convert_box_vertices"): "(center_x, center_y, center_z, l, w, h, yaw)",
        (box, "_estimate_yaw_pca"): "(rotated_pc)",
        (box, "_estimate_yaw_convex_hull"): "(rotated_pc)",
        (box, "save_3d_with_ground_alignment_bbox"): "(scene_dir, bbox_method='pca')",
        (cam, "length"): "(x, eps=1e-20)",
        (cam, "safe_normalize"): "(x, eps=1e-20)",
        (cam, "look_at"): "(campos, target, opengl=True)",
        (cam, "orbit_camera"): "(elevation, azimuth, radius=1, is_degree=True, target=None, opengl=True)",
    }
    for (mod, name), sig in want.items():
        assert str(inspect.signature(getattr(mod, name))) == sig, name
    # estimate_bbox keeps the reference's four leading parameters and defaults; extras are keyword additions
    params = list(inspect.signature(box.estimate_bbox).parameters.values())
    assert [p.name for p in params[:4]] == ["in_pc", "cat_name", "ground_equ", "method"]
    assert [p.default for p in params[1:4]] == [None, None, "pca"]
