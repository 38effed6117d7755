"""Drop-in for the hot-path functions of the reference's ``src/util.py``:
``depth_to_points`` (``:52-75``), ``project_to_2d`` (``:227-229``), ``draw_cube`` (``:232-289``)
and, either side of the box fit, ``analyze_mask`` (``:291-326``), ``get_maximum_height``
(``:328-335``), ``read_bounding_boxes_segmentations`` (``:337-382``, with
``create_boolean_mask_from_polygon`` ``:386-415`` and ``replace_categories_with_supercategories``
``:452-461``) and ``align_to_depth_match`` (``:464-494``).  Same names, signatures and return
types; the arithmetic runs on the B200.

Only these are provided: the rest of the reference's ``util.py`` is model / IO glue outside this
path.
"""

from __future__ import annotations

import json
import os

import numpy as np
import torch

from labelany3d_b200 import ops as _ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("labelany3d_b200 needs a CUDA device: this path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def depth_to_points(depth, K=None, R=None, t=None):
    """``depth[bs,H,W]`` -> points ``[H,W,3]`` float64 of batch element 0, like the reference.

    ``K`` is required (the reference fails on ``np.linalg.inv(None)`` too).  The intrinsics are
    inverted on the host with ``np.linalg.inv`` exactly as the reference does and the kernel then
    evaluates the reference's operation order in float64, so for ``K``-only calls (the pipeline's
    use, ``src/batch_scripts/depth.py:154``) the result is bit-identical to NumPy's.
    """
    Kinv = np.linalg.inv(K)            # raises like the reference for K=None / singular K
    dev = _device()
    d = depth.detach() if isinstance(depth, torch.Tensor) else torch.as_tensor(np.asarray(depth))
    if d.dim() != 3:
        raise ValueError(f"expected depth of shape [bs,H,W], got {tuple(d.shape)}")
    d0 = d[:1].to(device=dev, dtype=torch.float32).contiguous()      # only element 0 is returned
    Rd = None if R is None else torch.as_tensor(np.asarray(R, dtype=np.float64).copy(), device=dev)
    td = None if t is None else torch.as_tensor(np.asarray(t, dtype=np.float64).copy(), device=dev)
    out = _ops.depth_lift(d0, torch.as_tensor(np.ascontiguousarray(Kinv, dtype=np.float64), device=dev), Rd, td,
                          out_dtype=torch.float64, k_is_inverse=True)
    return out[0].cpu().numpy()


def project_to_2d(point_3d, camera_matrix):
    """``[u,v] = (K p)[:2] / (K p)[2]`` for one point ``[3]`` (reference API) or many ``[n,3]``."""
    dev = _device()
    p = np.asarray(point_3d, dtype=np.float64)
    single = p.ndim == 1
    uv = _ops.project_points(torch.as_tensor(np.ascontiguousarray(p.reshape(-1, 3)), device=dev),
                             torch.as_tensor(np.ascontiguousarray(camera_matrix, dtype=np.float64), device=dev))
    uv = uv.cpu().numpy()
    return uv[0] if single else uv


def draw_cube(scene_dir, is_ground=False):
    """Draw the scene's boxes over ``input.png`` (``src/util.py:232-289``); drawing stays on the
    host with OpenCV like the reference, the corner projection of ALL boxes is one kernel launch."""
    import cv2
    from PIL import Image
    with open(os.path.join(scene_dir, "cam_params.json")) as f:
        K = np.array(json.load(f)["K"])
    with open(os.path.join(scene_dir, "3dbbox_ground.json" if is_ground else "3dbbox.json")) as f:
        cubes = json.load(f)
    image = cv2.cvtColor(np.array(Image.open(os.path.join(scene_dir, "input.png"))), cv2.COLOR_RGB2BGR)
    if cubes:
        corners = np.array([c["bbox3D_cam"] for c in cubes], dtype=np.float64).reshape(-1, 3)
        all_uv = project_to_2d(corners, K).reshape(len(cubes), 8, 2)
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    for cube, uv in zip(cubes, all_uv if cubes else []):
        top = int(np.argmin(uv[:, 1]))          # first minimum, like the reference's scan
        for p in uv:
            cv2.circle(image, tuple(np.round(p).astype(int)), radius=3, color=(0, 255, 0), thickness=-1)
        for a, b in edges:
            cv2.line(image, tuple(np.round(uv[a]).astype(int)), tuple(np.round(uv[b]).astype(int)), (255, 0, 0), 2)
        cv2.putText(image, f'{cube["category_name"]}', (int(uv[top][0]), int(uv[top][1]) - 10),
                    cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 255), 1)
    cv2.imwrite(os.path.join(scene_dir, "vis_3dbox.png" if is_ground else "vis_3dbox_no_ground.png"), image)


# ---------------------------------------------------------------------------------------------
# mask statistics (scope table row f1)
# ---------------------------------------------------------------------------------------------
def _plane_stats(mask, boundary_threshold=10):
    dev = _device()
    m = np.asarray(mask)
    if m.ndim != 2:
        raise ValueError(f"expected a mask of shape [H,W], got {m.shape}")
    t = torch.as_tensor(np.ascontiguousarray(m != 0), device=dev)
    bits, _ = _ops.mask_scan(t[None])
    return _ops.mask_stats(bits, m.shape[0], m.shape[1], boundary_threshold)[0].cpu().numpy()


def analyze_mask(mask, image_size, scale_threshold=100, boundary_threshold=10):
    """``(is_truncated, is_scaleable)``: does the mask put at least 10 pixels into the border bands,
    and does it cover at least ``scale_threshold`` pixels (``src/util.py:291-326``).  The counts
    come from one pass over the mask on the GPU (``la3d_mask_scan`` + ``la3d_mask_stats``)."""
    mask = np.asarray(mask)
    if not np.array_equal(mask, mask.astype(bool)):
        raise ValueError("Image Mask must be binary (contain only 0s and 1s).")
    s = _plane_stats(mask, boundary_threshold)
    total_truncation = int(s[_ops.STAT_TOP]) + int(s[_ops.STAT_BOTTOM]) + int(s[_ops.STAT_LEFT]) + int(s[_ops.STAT_RIGHT])
    return total_truncation >= 10, int(s[_ops.STAT_AREA]) >= scale_threshold


def get_maximum_height(binary_mask):
    """Last minus first non-empty row plus one, 0 for an empty mask (``src/util.py:328-335``)."""
    s = _plane_stats(binary_mask)
    if s[_ops.STAT_ROWS] == 0:
        return 0
    return np.int64(int(s[_ops.STAT_LAST_ROW]) - int(s[_ops.STAT_FIRST_ROW]) + 1)


# ---------------------------------------------------------------------------------------------
# annotations -> mask stack (the loader in front of the path)
# ---------------------------------------------------------------------------------------------
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "coco_category_names.json")) as _f:
    COCO_CATEGORIES = {int(k): v for k, v in json.load(_f).items()}       # id -> name, src/util.py:419-449


def replace_categories_with_supercategories(category_ids, json_file_path=None):
    """Category names of COCO / COCONUT ids, ``"unknown"`` for anything else (``src/util.py:452-461``)."""
    return [COCO_CATEGORIES.get(c, "unknown") for c in category_ids]


def create_boolean_mask_from_polygon(image_shape, segmentation):
    """``(bool mask [image_shape[1], image_shape[0]], get_maximum_height(mask))`` of a polygon list or an
    RLE dict (``src/util.py:386-415``).  Polygons are filled on the host by OpenCV like the reference
    (its scan conversion is the format's definition); RLE runs are decoded on the GPU."""
    import cv2
    mask = np.zeros((image_shape[1], image_shape[0]), dtype=np.uint8)
    if isinstance(segmentation, list):
        for polygon in segmentation:
            points = np.array(polygon).reshape(-1, 2).astype(np.int32)
            cv2.fillPoly(mask, [points], color=1)
    elif isinstance(segmentation, dict):
        from labelany3d_b200 import coco_rle
        runs, (h, w) = coco_rle.runs_of(segmentation)
        counts, offsets, max_runs = coco_rle.pack_runs([runs])
        dev = _device()
        bits, _, status = _ops.rle_decode(torch.as_tensor(counts.view(np.int32), device=dev),
                                          torch.as_tensor(offsets, device=dev), h, w, max_runs)
        if int(status[0]) != 0:
            raise ValueError("Invalid RLE mask representation")
        mask = np.maximum(mask, _ops.unpack_bits(bits, h, w)[0].astype(np.uint8))
    boolean_mask = mask.astype(bool)
    return boolean_mask, get_maximum_height(boolean_mask)


def read_bounding_boxes_segmentations(annotations_path_or_list, image_size):
    """``(bboxes, masks bool [I,H,W], arange(I), category names)`` of one image's COCO / COCONUT
    annotations (``src/util.py:337-382``): crowd annotations are skipped; an instance is kept when its
    mask spans more than 6.25 % of the image height, puts fewer than 10 pixels into the 10-pixel border
    bands and covers at least 100 pixels.

    All run-length annotations of the image are decoded in ONE launch straight into bit planes
    (``la3d_rle_decode``), polygon masks (filled by OpenCV on the host like the reference) are scanned
    into the same layout, and one ``la3d_mask_stats`` launch yields every integer the admission test
    needs; only the kept planes come back to the host.  ``image_size = (width, height)``."""
    from labelany3d_b200 import coco_rle
    if isinstance(annotations_path_or_list, (str, os.PathLike)):
        with open(annotations_path_or_list, "r") as file:
            annotations = json.load(file)
    else:
        annotations = annotations_path_or_list

    # what each annotation is: ("crowd",) | ("none",) | ("rle", runs, (h, w)) | ("poly", mask, height)
    kinds = []
    for annotation in annotations:
        if annotation["iscrowd"]:
            kinds.append(("crowd",))
        elif "segmentation" not in annotation:
            kinds.append(("none",))
        else:
            seg = annotation["segmentation"]
            if isinstance(seg, dict) and "counts" in seg:
                runs, size = coco_rle.runs_of(seg)
                kinds.append(("rle", runs, size))
            elif isinstance(seg, list):
                mask = np.zeros((image_size[1], image_size[0]), dtype=np.uint8)
                import cv2
                for polygon in seg:
                    cv2.fillPoly(mask, [np.array(polygon).reshape(-1, 2).astype(np.int32)], color=1)
                kinds.append(("poly", mask.astype(bool)))
            else:
                raise KeyError("counts")            # a dict without run lengths: the reference fails on rle['counts']

    dev = _device() if any(k[0] in ("rle", "poly") for k in kinds) else None
    stats = {}                                     # annotation index -> the 8 integers of la3d_mask_stats
    planes = {}                                    # annotation index -> (bits tensor, row, (h, w)) for RLE planes
    by_size = {}
    for idx, k in enumerate(kinds):
        if k[0] == "rle":
            by_size.setdefault(k[2], []).append(idx)
    for (h, w), members in by_size.items():
        counts, offsets, max_runs = coco_rle.pack_runs([kinds[i][1] for i in members])
        bits, _, status = _ops.rle_decode(torch.as_tensor(counts.view(np.int32), device=dev),
                                          torch.as_tensor(offsets, device=dev), h, w, max_runs)
        if bool((status != 0).any()):
            raise ValueError("Invalid RLE mask representation")
        st = _ops.mask_stats(bits, h, w, 10).cpu().numpy().astype(np.int64)
        for row, i in enumerate(members):
            stats[i] = st[row]
            planes[i] = (bits, row, (h, w))
    poly = [i for i, k in enumerate(kinds) if k[0] == "poly"]
    if poly:
        stack = torch.as_tensor(np.ascontiguousarray(np.stack([kinds[i][1] for i in poly])), device=dev)
        pbits, _ = _ops.mask_scan(stack)
        st = _ops.mask_stats(pbits, image_size[1], image_size[0], 10).cpu().numpy().astype(np.int64)
        for row, i in enumerate(poly):
            stats[i] = st[row]

    bboxes, kept, category_ids = [], [], []
    for idx, (annotation, k) in enumerate(zip(annotations, kinds)):
        if k[0] == "crowd":
            print("Skip crowd annotation")
            continue
        if k[0] == "none":
            continue
        s = stats[idx]
        if k[0] == "rle":
            height = s[_ops.STAT_ROWS]                                   # np.sum(np.any(mask, axis=1)), :369-370
        else:
            height = 0 if s[_ops.STAT_ROWS] == 0 else s[_ops.STAT_LAST_ROW] - s[_ops.STAT_FIRST_ROW] + 1   # :328-335
        is_truncated = s[_ops.STAT_TOP] + s[_ops.STAT_BOTTOM] + s[_ops.STAT_LEFT] + s[_ops.STAT_RIGHT] >= 10
        is_scaleable = s[_ops.STAT_AREA] >= 100
        if height / image_size[1] > 0.0625 and not is_truncated and is_scaleable:
            kept.append(idx)
            category_ids.append(annotation["category_id"])
            bboxes.append(annotation["bbox"])
        else:
            print("Too small segmentation")

    segmentation_mask = []
    fetched = {}                                   # one device->host copy per decode launch, kept rows only
    for idx in kept:
        if kinds[idx][0] == "poly":
            segmentation_mask.append(kinds[idx][1])
        else:
            bits, row, (h, w) = planes[idx]
            key = id(bits)
            if key not in fetched:
                rows = [planes[j][1] for j in kept if kinds[j][0] == "rle" and planes[j][0] is bits]
                sel = torch.as_tensor(rows, device=bits.device)
                fetched[key] = dict(zip(rows, _ops.unpack_bits(bits.index_select(0, sel), h, w)))
            segmentation_mask.append(fetched[key][row])
    return (bboxes, np.array(segmentation_mask), np.arange(len(segmentation_mask)),
            replace_categories_with_supercategories(category_ids))


# ---------------------------------------------------------------------------------------------
# depth-scale alignment (scope table row f3)
# ---------------------------------------------------------------------------------------------
def align_to_depth_match(mask, depth_map, object_name, project_root, model):
    """4x4 transform that rescales a reconstructed object to the scene depth
    (``src/util.py:464-494``): the matcher the reference calls (``matching.process_image_space``,
    outside this path) renders the object; the scale is the median of ``depth_map / rendered
    depth`` over ``mask & rendered alpha > 0``, computed on the GPU straight from the bit planes
    (``la3d_masked_ratio_median``, bit-exact with ``np.median`` on float32)."""
    from matching.process_image_space import process_object
    R, T, image_render, depth = process_object(object_name, project_root, model)
    render_mask = image_render[..., -1] > 0
    dev = _device()
    mask = np.asarray(mask)
    H, W = mask.shape
    both = torch.as_tensor(np.ascontiguousarray(np.stack([mask != 0, render_mask])), device=dev)
    bits, _ = _ops.mask_scan(both)
    dm = np.asarray(depth_map)
    dr = np.asarray(depth)
    if dm.dtype != np.float32 or dr.dtype != np.float32:
        raise TypeError("align_to_depth_match: the GPU path takes float32 depth maps (what the pipeline stores)")
    n, scale = _ops.depth_scale_median(torch.as_tensor(np.ascontiguousarray(dm), device=dev)[None],
                                       torch.as_tensor(np.ascontiguousarray(dr), device=dev)[None, None],
                                       bits[0:1], bits[1:2], H, W)
    if int(n.item()) == 0:
        print("No overlap between masks found")
        return np.eye(4)
    scale = scale.cpu().numpy()[0, 0]
    transform = np.eye(4)
    transform[:3, :3] = np.linalg.inv(R[:3, :3]) * scale
    transform[:3, -1] = T[:3] * scale
    return transform
