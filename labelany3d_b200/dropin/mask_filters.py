"""Drop-in for ``filter_component_masks`` of the reference's ``src/model_wrappers.py:33-37`` (the one
function of that module on this path; patch it in with
``model_wrappers.filter_component_masks = mask_filters.filter_component_masks``) and the batched
admission test of ``read_bounding_boxes_segmentations`` (``src/util.py:369-375``)."""

from __future__ import annotations

import numpy as np
import torch

from labelany3d_b200 import ops as _ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("labelany3d_b200 needs a CUDA device: this path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def filter_component_masks(masks, foreground_mask, threshold=0.5):
    """Indices of the masks whose overlap with ``foreground_mask`` exceeds ``threshold`` of their own
    area, and of the rest.  The two pixel counts per mask are exact integers from the GPU
    (``la3d_mask_scan`` + ``la3d_mask_overlap``); the ratio test is the reference's float64 expression."""
    masks = np.asarray(masks)
    all_instances = np.arange(len(masks))
    if len(masks) == 0:
        return all_instances, all_instances
    dev = _device()
    H, W = masks.shape[-2:]
    fg = np.broadcast_to(np.asarray(foreground_mask), (H, W))
    stack = torch.as_tensor(np.ascontiguousarray(np.concatenate([masks != 0, (fg != 0)[None]])), device=dev)
    bits, _ = _ops.mask_scan(stack)
    I = len(masks)
    inter = _ops.mask_overlap(bits[:I], bits[I:], H, W, group=I).cpu().numpy().astype(np.int64)
    area = _ops.mask_stats(bits[:I], H, W)[:, _ops.STAT_AREA].cpu().numpy().astype(np.int64)
    is_foreground = (inter + 1e-6) / (area + 1e-6) > threshold
    return all_instances[is_foreground], all_instances[~is_foreground]


def admissible_instances(masks, image_size):
    """For a mask stack ``[I,H,W]``: which instances pass ``height/image_height > 0.0625 and not
    is_truncated and is_scaleable`` (``src/util.py:369-375``; ``image_size = (width, height)``,
    default thresholds of ``analyze_mask``).  One scan of the stack instead of ``I`` NumPy passes."""
    masks = np.asarray(masks)
    dev = _device()
    H, W = masks.shape[-2:]
    bits, _ = _ops.mask_scan(torch.as_tensor(np.ascontiguousarray(masks != 0), device=dev))
    s = _ops.mask_stats(bits, H, W, 10).cpu().numpy().astype(np.int64)
    truncated = s[:, _ops.STAT_TOP] + s[:, _ops.STAT_BOTTOM] + s[:, _ops.STAT_LEFT] + s[:, _ops.STAT_RIGHT] >= 10
    scaleable = s[:, _ops.STAT_AREA] >= 100
    return (s[:, _ops.STAT_ROWS] / image_size[1] > 0.0625) & ~truncated & scaleable
