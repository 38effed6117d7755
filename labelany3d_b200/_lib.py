"""ctypes binding of ``libla3d_sm100a.so`` (the C ABI declared in ``include/la3d.h``).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""

from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_vp, _i, _u32, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t

# name -> (restype, argtypes); mirrors include/la3d.h one to one
SIGNATURES = {
    "la3d_version": (_i, []),
    "la3d_last_error": (C.c_char_p, []),
    "la3d_depth_lift": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "la3d_chunks_per_plane": (_sz, [_i, _i]),
    "la3d_words_per_plane": (_sz, [_i, _i]),
    "la3d_mask_scan": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "la3d_mask_scan_thin": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "la3d_mask_stats": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "la3d_mask_overlap": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "la3d_prep_bytes": (_sz, [_i, _i]),
    "la3d_fit_prepare": (_i, [_vp, _vp, _i, _i, _u32, _u32, _vp, _sz, _vp]),
    "la3d_set_mt_blocks": (None, [_i]),
    "la3d_set_sample_seg_blocks": (None, [_i]),
    "la3d_sample_ranks": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "la3d_fit_scanned": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "la3d_fit_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "la3d_fit_boxes": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _u32, _u32, _vp, _sz, _vp, _i, _vp]),
    "la3d_fit_boxes_to": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _u32, _u32, _vp, _sz, _vp, _vp]),
    "la3d_fit_scanned_to": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "la3d_fit_boxes_rle_to": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u32, _u32, _vp, _sz, _vp, _vp, _vp]),
    "la3d_fit_all_points_to": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "la3d_fit_boxes_all_to": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "la3d_peer_signal": (_i, [_vp, _i, _i, _u32, _vp]),
    "la3d_peer_wait": (_i, [_vp, _i, _i, _u32, _vp, _vp]),
    "la3d_peer_barrier": (_i, [_vp, _i, _i, _u32, _vp, _vp]),
    "la3d_set_peer_timeout_ms": (None, [C.c_longlong]),
    "la3d_set_pipeline_images": (None, [_i]),
    "la3d_debug_fit_clocks": (None, [_vp]),
    "la3d_debug_sample_clocks": (None, [_vp]),
    "la3d_debug_step_events": (None, [_vp]),
    "la3d_debug_scatter_read": (_i, [_vp, _sz, _sz, _i, _i, _i, _i, _i, _vp, _vp]),
    "la3d_fit_points": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "la3d_iou_matrix": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "la3d_box2d_from_corners": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "la3d_masked_ratio_median": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "la3d_rle_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "la3d_fit_bits_workspace_bytes": (_sz, [_i, _i]),
    "la3d_fit_boxes_bits": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u32, _u32, _vp, _sz, _vp, _i, _vp]),
    "la3d_fit_boxes_rle": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u32, _u32, _vp, _sz, _vp, _vp, _i, _vp]),
    "la3d_fit_all_points": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "la3d_fit_boxes_all": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp, _i, _vp]),
    "la3d_ransac_subset_fit": (_i, [_vp, _vp, _vp, C.c_longlong, _vp, _vp]),
    "la3d_ransac_classify": (_i, [_vp, _vp, C.c_longlong, C.c_float, C.c_float, _vp, _vp]),
    "la3d_scale_fill": (_i, [_vp, _vp, C.c_longlong, C.c_float, C.c_float, _vp, _vp]),
    "la3d_project_points": (_i, [_vp, _vp, _vp, C.c_longlong, _vp, _vp]),
}

MAX_PEERS = 8


class Sink(C.Structure):
    """``la3d_sink`` of include/la3d.h: the destinations of a fit's records and the peer synchronisation."""
    _fields_ = [("records", _vp * MAX_PEERS), ("flags", _vp * MAX_PEERS), ("counter", _vp), ("status", _vp),
                ("epoch", _u32), ("n_out", C.c_int32), ("rank", C.c_int32), ("rec_f64", C.c_int32)]


def make_sink(records, rec_f64, flags=None, counter=None, status=None, epoch=0, rank=0):
    """``records``: list of device pointers (ints); ``flags``: list of the ranks' flag-row pointers or None."""
    s = Sink()
    for p, ptr in enumerate(records):
        s.records[p] = ptr
    s.n_out = len(records)
    s.rec_f64 = int(bool(rec_f64))
    if flags is not None:
        for p, ptr in enumerate(flags):
            s.flags[p] = ptr
        s.counter, s.status, s.epoch, s.rank = counter, status, int(epoch) & 0xFFFFFFFF, int(rank)
    return s


_lib = None


class La3dError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.isfile(path):
        raise La3dError(
            f"{path} is missing: build it with `python -m labelany3d_b200.build` "
            "(needs nvcc; there is no CPU fallback for this path)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().la3d_last_error().decode("utf-8", "replace")
        raise La3dError(f"{what} failed with code {rc}: {msg}")
