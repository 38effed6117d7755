"""CPU oracle for the "next" rows of the scope table.  TEST INFRASTRUCTURE ONLY.

NumPy restatements of the reference functions either side of the box-fitting path.  Like
``la3d_oracle`` this module is a checker: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it; nothing under ``labelany3d_b200/`` does.

Parity status: PINNED against the unmodified reference functions, imported live by
``tests/golden/make_golden_next.py`` (which also writes ``tests/golden/golden_next_v1.npz``).

=================================  ====================================================
oracle function                    reference (paths relative to the reference root)
=================================  ====================================================
``analyze_mask``                   ``src/util.py:291-326``
``get_maximum_height``             ``src/util.py:328-335``
``rows_with_pixels``               ``src/util.py:369-370`` (``height = np.sum(rows)``)
``keep_instance``                  ``src/util.py:374-375`` (the admission test)
``mask_stats``                     the integers behind the four functions above
``filter_component_masks``         ``src/model_wrappers.py:33-37``
``iou2D`` / ``iou_matrix``         ``src/tools/combine_results.py:111-123, 130-134``
``hungarian_matching``             ``src/tools/combine_results.py:126-144`` (SciPy
                                   ``linear_sum_assignment``, like the reference)
``box2d_proj_trunc``               ``src/tools/combine_results.py:234-252``
``depth_scale_median``             ``src/util.py:473-486`` (overlap, gathers, median ratio)
=================================  ====================================================
"""

from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ f1: mask statistics
def analyze_mask(mask, image_size, scale_threshold=100, boundary_threshold=10):
    """``(is_truncated, is_scaleable)`` of one mask (``src/util.py:291-326``)."""
    mask = np.asarray(mask)
    if not np.array_equal(mask, mask.astype(bool)):
        raise ValueError("Image Mask must be binary (contain only 0s and 1s).")
    b = boundary_threshold
    scale = np.sum(mask)
    top = np.sum(mask[:b, :])
    bottom = np.sum(mask[-b:, :])
    left = np.sum(mask[:, :b])
    right = np.sum(mask[:, -b:])
    return (top + bottom + left + right) >= 10, scale >= scale_threshold


def get_maximum_height(binary_mask):
    rows = np.where(np.any(binary_mask, axis=1))[0]
    if rows.size == 0:
        return 0
    return rows[-1] - rows[0] + 1


def rows_with_pixels(mask):
    return np.sum(np.any(mask, axis=1))


def keep_instance(mask, image_size):
    """The admission test of ``read_bounding_boxes_segmentations`` (``src/util.py:373-375``);
    ``image_size`` is ``(width, height)``."""
    height = rows_with_pixels(mask)
    is_truncated, is_scaleable = analyze_mask(mask, image_size)
    return bool(height / image_size[1] > 0.0625 and not is_truncated and is_scaleable)


def mask_stats(masks, boundary_threshold=10):
    """``[P, 8]`` int32 per plane: area, top, bottom, left, right band counts, first / last
    non-empty row (-1 if empty), number of non-empty rows."""
    m = np.asarray(masks).astype(bool)
    m = m.reshape((-1,) + m.shape[-2:])
    b = boundary_threshold
    out = np.zeros((m.shape[0], 8), dtype=np.int32)
    for p, plane in enumerate(m):
        rows = np.where(plane.any(axis=1))[0]
        out[p] = (plane.sum(), plane[:b, :].sum(), plane[-b:, :].sum(), plane[:, :b].sum(), plane[:, -b:].sum(),
                  rows[0] if rows.size else -1, rows[-1] if rows.size else -1, rows.size)
    return out


def filter_component_masks(masks, foreground_mask, threshold=0.5):
    """``src/model_wrappers.py:33-37``: indices of the masks whose overlap with the foreground
    exceeds ``threshold`` of their area, and of the others."""
    all_instances = np.arange(len(masks))
    is_foreground = ((masks & foreground_mask).sum((-1, -2)) + 1e-6) / (masks.sum((-1, -2)) + 1e-6) > threshold
    return all_instances[is_foreground], all_instances[~is_foreground]


# ------------------------------------------------------------------ f2: combine_results arithmetic
def iou2D(box1, box2):
    x1 = max(box1[0], box2[0])
    y1 = max(box1[1], box2[1])
    x2 = min(box1[2], box2[2])
    y2 = min(box1[3], box2[3])
    intersection = max(0, x2 - x1) * max(0, y2 - y1)
    area1 = (box1[2] - box1[0]) * (box1[3] - box1[1])
    area2 = (box2[2] - box2[0]) * (box2[3] - box2[1])
    return intersection / (area1 + area2 - intersection + 1e-6)


def iou_matrix(boxes0, boxes1):
    out = np.zeros((len(boxes0), len(boxes1)))
    for i, b0 in enumerate(boxes0):
        for j, b1 in enumerate(boxes1):
            out[i, j] = iou2D(b0, b1)
    return out


def hungarian_matching(boxes0, boxes1):
    from scipy.optimize import linear_sum_assignment
    cost = -iou_matrix(boxes0, boxes1)
    rows, cols = linear_sum_assignment(cost)
    return [(i, j, -cost[i, j]) for i, j in zip(rows, cols)]


def box2d_proj_trunc(corners, K, W, H):
    """``bbox2D_proj`` and ``bbox2D_trunc`` of one box from its 8 corners
    (``src/tools/combine_results.py:234-252``; Python ``min`` / ``max`` semantics)."""
    pts = [np.dot(K, np.array(p))[:2] / np.dot(K, np.array(p))[2] for p in np.asarray(corners)]
    min_x = min(p[0] for p in pts)
    min_y = min(p[1] for p in pts)
    max_x = max(p[0] for p in pts)
    max_y = max(p[1] for p in pts)
    return [min_x, min_y, max_x, max_y], [max(0, min_x), max(0, min_y), min(W, max_x), min(H, max_y)]


# ------------------------------------------------------------------ f3: depth-scale alignment
def depth_scale_median(mask, depth_map, render_mask, depth_render):
    """``(n_overlap, scale)`` of ``align_to_depth_match`` (``src/util.py:473-486``): the median of
    ``depth_map / depth_render`` over ``mask & render_mask``; ``scale`` is ``None`` when the
    overlap is empty (the reference returns the identity transform then)."""
    overlap = mask & render_mask
    if not overlap.any():
        return 0, None
    ratios = depth_map[overlap] / depth_render[overlap]
    return int(overlap.sum()), np.median(ratios)
