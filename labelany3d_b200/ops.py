"""Host side of the B200 box-fitting path: torch tensors in, torch tensors out.

Every function here drives the hand-written sm_100a kernels of
``libla3d_sm100a.so`` through the C ABI (``include/la3d.h``) with raw device
pointers on the current CUDA stream.  torch is used for device memory and
streams only.  Inputs must live on a CUDA device; nothing falls back to the CPU.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .records import METHODS, REC, SUBSAMPLE


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(name, t, dtype=None, pinned_ok=False):
    """``pinned_ok``: page-locked host memory is accepted too (the kernels read it in place over PCIe
    through unified addressing; the arithmetic still runs on the GPU)."""
    if not isinstance(t, torch.Tensor) or not (t.is_cuda or (pinned_ok and t.is_pinned())):
        raise TypeError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _masks_u8(masks):
    """bool -> (uint8 view, is_01=1); uint8 -> (as is, 0: any nonzero byte is 'set')."""
    if masks.dtype == torch.bool:
        return masks.view(torch.uint8), 1
    if masks.dtype == torch.uint8:
        return masks, 0
    raise TypeError(f"masks must be bool or uint8, got {masks.dtype}")


def _method_id(method):
    # unknown names are passed through as an invalid id so the kernel reports
    # LA3D_ST_BAD_METHOD per box, the way the reference raises per call
    return METHODS.get(method, 99)


# ---------------------------------------------------------------------------------------
def depth_lift(depth, K, R=None, t=None, out_dtype=torch.float32, k_is_inverse=False):
    """``depth[B,H,W]`` float32 -> camera-space points ``[B,H,W,3]``.

    Batched form of the reference's ``depth_to_points`` (``src/util.py:52-75``).
    ``K`` is ``[B,3,3]`` or one shared ``[3,3]`` (float64).  ``out_dtype=float64``
    reproduces the reference's operation order bit for bit; ``float32`` is the
    bandwidth-lean form (16 B per pixel), rounded once from the float64 value.
    """
    lib = _lib.load()
    depth = _need_cuda("depth", depth, torch.float32)
    B, H, W = depth.shape
    K = _need_cuda("K", K, torch.float64)
    if K.shape == (3, 3):
        k_stride = 0
    elif K.shape == (B, 3, 3):
        k_stride = 9
    else:
        raise ValueError(f"K must be [3,3] or [{B},3,3], got {tuple(K.shape)}")
    if R is not None:
        R = _need_cuda("R", R, torch.float64)
        assert R.shape == (3, 3)
    if t is not None:
        t = _need_cuda("t", t, torch.float64)
        assert t.shape == (3,)
    if out_dtype not in (torch.float32, torch.float64):
        raise TypeError("out_dtype must be float32 or float64")
    out = torch.empty((B, H, W, 3), dtype=out_dtype, device=depth.device)
    with torch.cuda.device(depth.device):
        rc = lib.la3d_depth_lift(_ptr(depth), _ptr(K), k_stride, int(bool(k_is_inverse)), _ptr(R), _ptr(t), B, H, W,
                                 _ptr(out), int(out_dtype == torch.float64), _stream())
    _lib.check(rc, "la3d_depth_lift")
    return out


def scan_layout(H, W):
    """(chunks per plane, bit-words per plane) of the mask scan for an ``H x W`` image."""
    lib = _lib.load()
    return int(lib.la3d_chunks_per_plane(H, W)), int(lib.la3d_words_per_plane(H, W))


def mask_scan(masks, thin=None):
    """``masks[..., H, W]`` (bool / uint8) -> ``(bits[P, words] int32, chunk_counts[P, chunks] int32)``.

    ``thin=(ctas_per_sm, stages)`` runs the persistent TMA-fed form (``la3d_mask_scan_thin``; needs
    ``H*W % 512 == 0``) instead of the one-tile-per-CTA kernel; the results are identical.

    ``P`` = product of the leading dimensions.  Bit ``k`` of word ``w`` of a plane is
    pixel ``32 w + k`` in row-major order; a ``chunk_counts`` word packs the set-pixel counts
    of the four 128-pixel quarters of a 512-pixel chunk, one byte each (byte 0 = first quarter).
    """
    lib = _lib.load()
    masks = _need_cuda("masks", masks)
    m8, is01 = _masks_u8(masks)
    H, W = masks.shape[-2:]
    planes = masks.numel() // (H * W)
    chunks, words = scan_layout(H, W)
    bits = torch.empty((planes, words), dtype=torch.int32, device=masks.device)
    cc = torch.empty((planes, chunks), dtype=torch.int32, device=masks.device)
    with torch.cuda.device(masks.device):
        if thin is None:
            rc = lib.la3d_mask_scan(_ptr(m8), planes, H, W, is01, _ptr(bits), _ptr(cc), _stream())
        else:
            rc = lib.la3d_mask_scan_thin(_ptr(m8), planes, H, W, is01, _ptr(bits), _ptr(cc), int(thin[0]), int(thin[1]),
                                         _stream())
    _lib.check(rc, "la3d_mask_scan")
    return bits, cc


def rle_decode(run_counts, run_offsets, H, W, max_runs=0):
    """COCO run-length planes -> ``(bits, chunk_counts, status)`` in the layout of :func:`mask_scan`,
    without the byte masks (``la3d_rle_decode``; replaces ``mask_utils.decode`` of
    ``src/util.py:361-370``).  ``run_counts`` uint32-valued int32/uint32 CUDA tensor of all planes'
    column-major run lengths, ``run_offsets[P+1]`` int64 CUDA tensor, ``max_runs`` the largest run
    count of one plane if the caller knows it (``coco_rle.pack_runs`` returns all three).  ``status[P]``
    int32: 0 ok, 1 the runs overflow the image (the in-image part is decoded), 2 plane refused."""
    lib = _lib.load()
    run_offsets = _need_cuda("run_offsets", run_offsets, torch.int64)
    run_counts = _need_cuda("run_counts", run_counts)
    if run_counts.dtype not in (torch.int32, torch.uint32):
        raise TypeError(f"run_counts must be int32 / uint32 (holding unsigned 32-bit values), got {run_counts.dtype}")
    planes = run_offsets.numel() - 1
    if planes <= 0:
        raise ValueError("run_offsets must hold at least two entries")
    chunks, words = scan_layout(H, W)
    dev = run_offsets.device
    bits = torch.empty((planes, words), dtype=torch.int32, device=dev)
    cc = torch.empty((planes, chunks), dtype=torch.int32, device=dev)
    status = torch.empty((planes,), dtype=torch.int32, device=dev)
    # scratch for planes whose run ends do not fit the shared-memory copy announced by max_runs
    ends_ws = torch.empty((max(run_counts.numel(), 1),), dtype=torch.int32, device=dev)
    counts_ptr = _ptr(run_counts) if run_counts.numel() else _ptr(ends_ws)      # never dereferenced when empty
    with torch.cuda.device(dev):
        rc = lib.la3d_rle_decode(counts_ptr, _ptr(run_offsets), planes, H, W, int(max_runs), _ptr(ends_ws), _ptr(bits),
                                 _ptr(cc), _ptr(status), _stream())
    _lib.check(rc, "la3d_rle_decode")
    return bits, cc, status


def unpack_bits(bits, H, W):
    """Bit planes ``[P, words]`` -> NumPy ``bool [P,H,W]`` on the host (1/8 of the bytes cross the bus)."""
    host = bits.detach().cpu().numpy().view(np.uint8)
    P = host.shape[0]
    return np.unpackbits(host, axis=1, bitorder="little")[:, :H * W].reshape(P, H, W).astype(bool)


STAT_AREA, STAT_TOP, STAT_BOTTOM, STAT_LEFT, STAT_RIGHT, STAT_FIRST_ROW, STAT_LAST_ROW, STAT_ROWS = range(8)


def mask_stats(bits, H, W, boundary_threshold=10):
    """Per-plane integer statistics from the bit planes of :func:`mask_scan` (``la3d_mask_stats``).

    Returns ``stats[P, 8]`` int32: area, pixels inside the top / bottom / left / right border band
    (``analyze_mask``, ``src/util.py:303-320`` of the reference: bands ``[:b]`` and ``[-b:]`` with
    Python slice semantics, so ``b = 0`` makes the bottom / right band the whole image), first and
    last non-empty row (-1 if the mask is empty), number of non-empty rows.
    """
    lib = _lib.load()
    bits = _need_cuda("bits", bits, torch.int32)
    planes = bits.shape[0]
    b = int(boundary_threshold)
    t0, t1, _ = slice(None, b).indices(H)
    b0, b1, _ = slice(-b, None).indices(H)
    l0, l1, _ = slice(None, b).indices(W)
    r0, r1, _ = slice(-b, None).indices(W)
    bands = (ctypes.c_int * 8)(t0, max(t1, t0), b0, max(b1, b0), l0, max(l1, l0), r0, max(r1, r0))
    stats = torch.empty((planes, 8), dtype=torch.int32, device=bits.device)
    with torch.cuda.device(bits.device):
        rc = lib.la3d_mask_stats(_ptr(bits), planes, H, W, bands, _ptr(stats), _stream())
    _lib.check(rc, "la3d_mask_stats")
    return stats


def mask_overlap(bits, other_bits, H, W, group=1):
    """``inter[p] = |bits[p] & other_bits[p // group]|`` (``la3d_mask_overlap``): the numerator of
    ``filter_component_masks`` (``src/model_wrappers.py:36`` of the reference)."""
    lib = _lib.load()
    bits = _need_cuda("bits", bits, torch.int32)
    other_bits = _need_cuda("other_bits", other_bits, torch.int32)
    planes = bits.shape[0]
    if other_bits.shape[0] * group < planes or other_bits.shape[1] != bits.shape[1]:
        raise ValueError("other_bits must hold one plane per group of `group` planes, same image size")
    inter = torch.empty((planes,), dtype=torch.int32, device=bits.device)
    with torch.cuda.device(bits.device):
        rc = lib.la3d_mask_overlap(_ptr(bits), _ptr(other_bits), planes, int(group), H, W, _ptr(inter), _stream())
    _lib.check(rc, "la3d_mask_overlap")
    return inter


def fit_prepare(K, ground, B, I, seed=0, image_offset=0):
    """Mask-independent preparation of a batch (``la3d_fit_prepare``): per-image MT19937 words,
    intrinsics and their inverse, per-box ground rotations.  Returns the opaque ``prep`` buffer
    (uint8 CUDA tensor) that :func:`sample_ranks` and :func:`fit_scanned` consume."""
    lib = _lib.load()
    K = _need_cuda("K", K, torch.float64)
    if tuple(K.shape) != (B, 3, 3):
        raise ValueError(f"K must be [{B},3,3]")
    if ground is not None:
        ground = _need_cuda("ground", ground, torch.float64)
        if tuple(ground.shape) != (B, I, 3):
            raise ValueError(f"ground must be [{B},{I},3]")
    nbytes = int(lib.la3d_prep_bytes(B, I))
    prep = torch.empty(nbytes, dtype=torch.uint8, device=K.device)
    with torch.cuda.device(K.device):
        rc = lib.la3d_fit_prepare(_ptr(K), _ptr(ground), B, I, int(seed) & 0xFFFFFFFF, int(image_offset) & 0xFFFFFFFF,
                                  _ptr(prep), nbytes, _stream())
    _lib.check(rc, "la3d_fit_prepare")
    return prep


def sample_ranks(chunk_counts, B, I, H, W, seed=0, image_offset=0, prep=None):
    """Per-plane pixel counts and the reference's 500 random rows of ``pts[mask]``.

    Returns ``(counts[B,I] int32, ranks[B,I,500] int32)``; ``ranks`` of planes with at
    most 500 pixels are left at -1 (the reference keeps all their points).  ``prep``: the
    buffer of :func:`fit_prepare` (made here, with identity cameras, when not given).
    """
    lib = _lib.load()
    cc = _need_cuda("chunk_counts", chunk_counts, torch.int32)
    if prep is None:
        eye = torch.eye(3, dtype=torch.float64, device=cc.device).expand(B, 3, 3).contiguous()
        prep = fit_prepare(eye, None, B, I, seed, image_offset)
    counts = torch.empty((B, I), dtype=torch.int32, device=cc.device)
    ranks = torch.full((B, I, SUBSAMPLE), -1, dtype=torch.int32, device=cc.device)
    with torch.cuda.device(cc.device):
        rc = lib.la3d_sample_ranks(_ptr(cc), _ptr(prep), B, I, H, W, _ptr(counts), _ptr(ranks), _stream())
    _lib.check(rc, "la3d_sample_ranks")
    return counts, ranks


def fit_scanned(depth, prep, bits, chunk_counts, ranks, method="pca", yaw_steps=0, out_dtype=torch.float64):
    """The fit kernel alone (``la3d_fit_scanned``) on the outputs of :func:`mask_scan`,
    :func:`fit_prepare` and :func:`sample_ranks`; returns ``records[B,I,64]``."""
    lib = _lib.load()
    depth = _need_cuda("depth", depth, torch.float32)
    B, H, W = depth.shape
    I = ranks.shape[1]
    rec = torch.empty((B, I, REC), dtype=out_dtype, device=depth.device)
    with torch.cuda.device(depth.device):
        rc = lib.la3d_fit_scanned(_ptr(depth), _ptr(prep), _ptr(bits), _ptr(chunk_counts), _ptr(ranks), B, I, H, W,
                                  _method_id(method), int(yaw_steps), _ptr(rec), int(out_dtype == torch.float64),
                                  _stream())
    _lib.check(rc, "la3d_fit_scanned")
    return rec


class BoxFitter:
    """Reusable plan for ``fit_boxes`` on a fixed shape: owns the workspace and the output.

    One call = three launches on the current stream (mask scan with the mask-independent
    preparation riding in its grid, subsample ranks, fit); nothing is allocated and nothing
    synchronises.
    """

    def __init__(self, B, I, H, W, device="cuda", out_dtype=torch.float64):
        self.lib = _lib.load()
        self.shape = (int(B), int(I), int(H), int(W))
        self.device = torch.device(device)
        if out_dtype not in (torch.float32, torch.float64):
            raise TypeError("out_dtype must be float32 or float64")
        self.out_dtype = out_dtype
        self.ws_bytes = int(self.lib.la3d_fit_workspace_bytes(*self.shape))
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        assert self.workspace.data_ptr() % 256 == 0
        self.records = torch.empty((B, I, REC), dtype=out_dtype, device=self.device)

    def __call__(self, depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, image_offset=0, out=None,
                 events=None, sink=None):
        """Fit every (image, instance) box; returns ``records[B,I,64]`` (a buffer owned by the plan
        unless ``out`` is given).  ``events``: optional list of ``torch.cuda.Event``.  Four: recorded by the
        library around the step's own three launches (before, after the scan with its preparation CTAs, after
        the sampler, after the fit).  Five: the four kernels are issued one after the other through the
        step-wise entry points with the events recorded before, between and after them: prepare, scan,
        sample, fit.  (Per-kernel timing without a profiler.)  ``sink``: a ``_lib.Sink`` (``la3d_sink``): the fit kernel then writes every
        record to all its destinations, e.g. this rank's slot in the peer-mapped gathered buffer of every
        rank, and synchronises with the peers itself (``la3d_fit_boxes_to``) instead of writing ``out``."""
        B, I, H, W = self.shape
        # the fit gathers only 500 depth values per box, so the depth maps may stay in pinned host memory
        depth = _need_cuda("depth", depth, torch.float32, pinned_ok=True)
        K = _need_cuda("K", K, torch.float64)
        masks = _need_cuda("masks", masks)
        if tuple(depth.shape) != (B, H, W) or tuple(masks.shape) != (B, I, H, W) or tuple(K.shape) != (B, 3, 3):
            raise ValueError(f"shapes do not match the plan {self.shape}: depth {tuple(depth.shape)}, "
                             f"masks {tuple(masks.shape)}, K {tuple(K.shape)}")
        if ground is not None:
            ground = _need_cuda("ground", ground, torch.float64)
            if tuple(ground.shape) != (B, I, 3):
                raise ValueError(f"ground must be [{B},{I},3]")
        m8, is01 = _masks_u8(masks)
        rec = self.records if out is None else _need_cuda("out", out, self.out_dtype)
        if tuple(rec.shape) != (B, I, REC):
            raise ValueError(f"out must be [{B},{I},{REC}]")
        f64 = int(self.out_dtype == torch.float64)
        lib = self.lib
        with torch.cuda.device(self.device):
            st = _stream()
            if sink is not None:
                if events is not None:
                    raise ValueError("events and sink are mutually exclusive")
                rc = lib.la3d_fit_boxes_to(_ptr(depth), _ptr(m8), _ptr(K), _ptr(ground), B, I, H, W, is01,
                                           _method_id(method), int(yaw_steps), int(seed) & 0xFFFFFFFF,
                                           int(image_offset) & 0xFFFFFFFF, _ptr(self.workspace), self.ws_bytes,
                                           ctypes.byref(sink), st)
                _lib.check(rc, "la3d_fit_boxes_to")
            elif events is None or len(events) == 4:
                if events is not None:
                    # the step's own three launches, timed from inside the call (la3d_debug_step_events)
                    for e in events:
                        e.record()                # torch creates the CUDA event lazily, on the first record
                    handles = (ctypes.c_void_p * 4)(*[e.cuda_event for e in events])
                    lib.la3d_debug_step_events(handles)
                try:
                    rc = self._fit_boxes(lib, depth, m8, K, ground, is01, method, yaw_steps, seed, image_offset, rec, f64, st)
                finally:
                    if events is not None:
                        lib.la3d_debug_step_events(None)
                _lib.check(rc, "la3d_fit_boxes")
            else:
                # the same kernels through the step-wise entry points, serialised, with events in between
                bits, cc, counts, ranks, prep, prep_bytes = self._carve()
                sd, off = int(seed) & 0xFFFFFFFF, int(image_offset) & 0xFFFFFFFF
                events[0].record()
                _lib.check(lib.la3d_fit_prepare(_ptr(K), _ptr(ground), B, I, sd, off, prep, prep_bytes, st), "la3d_fit_prepare")
                events[1].record()
                _lib.check(lib.la3d_mask_scan(_ptr(m8), B * I, H, W, is01, bits, cc, st), "la3d_mask_scan")
                events[2].record()
                _lib.check(lib.la3d_sample_ranks(cc, prep, B, I, H, W, counts, ranks, st), "la3d_sample_ranks")
                events[3].record()
                _lib.check(lib.la3d_fit_scanned(_ptr(depth), prep, bits, cc, ranks, B, I, H, W, _method_id(method),
                                                int(yaw_steps), _ptr(rec), f64, st), "la3d_fit_scanned")
                events[4].record()
        return rec

    def _fit_boxes(self, lib, depth, m8, K, ground, is01, method, yaw_steps, seed, image_offset, rec, f64, st):
        B, I, H, W = self.shape
        return lib.la3d_fit_boxes(_ptr(depth), _ptr(m8), _ptr(K), _ptr(ground), B, I, H, W, is01,
                                  _method_id(method), int(yaw_steps), int(seed) & 0xFFFFFFFF,
                                  int(image_offset) & 0xFFFFFFFF, _ptr(self.workspace), self.ws_bytes,
                                  _ptr(rec), f64, st)

    def capture(self, depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, image_offset=0, out=None):
        """Record one call as a CUDA graph and return its ``replay()``: the kernels of a step (and
        the prep CTAs riding in the scan) are then launched by one ``cudaGraphLaunch``
        instead of a dozen driver calls.  The graph reads the SAME buffers on every replay: refill
        ``depth`` / ``K`` / ``masks`` / ``ground`` in place to process new data.  Returns
        ``(replay, records)``."""
        rec = self.records if out is None else out
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                     # warm-up outside capture (lazy driver state)
                self(depth, K, masks, ground, method, yaw_steps, seed, image_offset, out=rec)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self(depth, K, masks, ground, method, yaw_steps, seed, image_offset, out=rec)
        self._graph_keepalive = (graph, depth, K, masks, ground, rec)
        return graph.replay, rec

    def _carve(self):
        """Workspace sub-buffers in the order ``la3d_fit_boxes`` lays them out (256-byte aligned)."""
        B, I, H, W = self.shape
        planes = B * I
        chunks, words = scan_layout(H, W)
        up = lambda v: (v + 255) & ~255  # noqa: E731
        base = self.workspace.data_ptr()
        o_cc = up(planes * words * 4)
        o_counts = up(o_cc + planes * chunks * 4)
        o_ranks = up(o_counts + planes * 4)
        o_prep = up(o_ranks + planes * SUBSAMPLE * 4)
        prep_bytes = int(self.lib.la3d_prep_bytes(B, I))
        assert up(o_prep + prep_bytes) == self.ws_bytes
        return base, base + o_cc, base + o_counts, base + o_ranks, base + o_prep, prep_bytes


class HostBoxFitter:
    """Boxes straight from page-locked HOST buffers: the end-to-end form of the path.

    Per call the mask stack, intrinsics and ground normals are copied to the device on a copy stream
    into one of two buffer sets, so the copy of batch k+1 overlaps the kernels of batch k; the depth
    maps are NOT copied: the fit reads 500 depth values per box, which the kernel gathers in place
    from the pinned host buffer over PCIe (a third of the input bytes never cross the bus); the records
    are copied back into a pinned host tensor.  ``fitter`` is a ``BoxFitter`` or a
    ``dist.ShardedBoxFitter`` (anything with their call signature).  Everything is asynchronous: a
    call returns an event that fires when its records have landed in ``out_host``; the host inputs
    of a call must stay untouched until then.
    """

    def __init__(self, fitter, B, I, H, W, device="cuda", depth_in_place=True):
        self.fitter = fitter
        self.device = torch.device(device)
        self.depth_in_place = bool(depth_in_place)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=self.device) for _ in range(2)]  # noqa: E731
        self.d_K, self.d_masks = mk((B, 3, 3), torch.float64), mk((B, I, H, W), torch.bool)
        self.d_ground = mk((B, I, 3), torch.float64)
        self.d_depth = None if self.depth_in_place else mk((B, H, W), torch.float32)
        self.ev_ready = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self.k = 0

    def h2d_bytes(self, B, I, H, W):
        """Bytes that cross the bus host->device per call (depth read in place: one 128-byte line request per
        sample, an upper bound - address-sorted samples of a box share lines)."""
        copied = B * 72 + B * I * H * W + B * I * 24
        return copied + (B * I * SUBSAMPLE * 128 if self.depth_in_place else B * H * W * 4)

    def __call__(self, depth, K, masks, ground, out_host, method="pca", yaw_steps=0, seed=0):
        for name, t in (("depth", depth), ("K", K), ("masks", masks), ("ground", ground), ("out_host", out_host)):
            if t is not None and not (isinstance(t, torch.Tensor) and t.is_pinned()):
                raise TypeError(f"{name} must be a pinned host tensor")
        s = self.k & 1
        self.k += 1
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_event(self.ev_free[s])          # the kernels of batch k-2 are done with this set
        with torch.cuda.stream(self.copy_stream):
            self.d_K[s].copy_(K, non_blocking=True)
            self.d_masks[s].copy_(masks, non_blocking=True)
            if ground is not None:
                self.d_ground[s].copy_(ground, non_blocking=True)
            if not self.depth_in_place:
                self.d_depth[s].copy_(depth, non_blocking=True)
            self.ev_ready[s].record(self.copy_stream)
        cur.wait_event(self.ev_ready[s])
        rec = self.fitter(depth if self.depth_in_place else self.d_depth[s], self.d_K[s], self.d_masks[s],
                          None if ground is None else self.d_ground[s], method, yaw_steps, seed=seed)
        out_host.copy_(rec, non_blocking=True)
        self.ev_free[s].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        return done


def fit_boxes(depth, K, masks, ground=None, method="pca", yaw_steps=0, seed=0, image_offset=0,
              out_dtype=torch.float64):
    """``depth[B,H,W], K[B,3,3], masks[B,I,H,W], ground[B,I,3]|None`` -> records ``[B,I,64]``.

    The composed path of SURVEY.md section 3.4 for a batch: per instance the
    reference's ``estimate_bbox`` applied to ``depth_to_points(depth)[mask]`` plus the
    2D reprojection of the corners, with the legacy NumPy RNG re-seeded to
    ``seed + image_offset + b`` at the start of image ``b``.  ``depth`` may stay in pinned host
    memory (read in place; the masks decide the device).
    """
    B, I, H, W = masks.shape
    return BoxFitter(B, I, H, W, device=masks.device, out_dtype=out_dtype)(
        depth, K, masks, ground, method, yaw_steps, seed, image_offset).clone()


def fit_boxes_bits(depth, K, bits, chunk_counts, I, ground=None, method="pca", yaw_steps=0, seed=0, image_offset=0,
                   out_dtype=torch.float64):
    """:func:`fit_boxes` from bit planes that already exist (``la3d_fit_boxes_bits``): the output of
    :func:`rle_decode` or of an earlier :func:`mask_scan`.  ``bits[B*I, words]``, ``chunk_counts[B*I, chunks]``."""
    lib = _lib.load()
    depth = _need_cuda("depth", depth, torch.float32, pinned_ok=True)
    K = _need_cuda("K", K, torch.float64)
    bits = _need_cuda("bits", bits, torch.int32)
    cc = _need_cuda("chunk_counts", chunk_counts, torch.int32)
    B, H, W = depth.shape
    chunks, words = scan_layout(H, W)
    if tuple(bits.shape) != (B * I, words) or tuple(cc.shape) != (B * I, chunks) or tuple(K.shape) != (B, 3, 3):
        raise ValueError("bit planes / chunk counts / K do not match depth and I")
    if ground is not None:
        ground = _need_cuda("ground", ground, torch.float64)
        if tuple(ground.shape) != (B, I, 3):
            raise ValueError(f"ground must be [{B},{I},3]")
    dev = bits.device
    ws_bytes = int(lib.la3d_fit_bits_workspace_bytes(B, I))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rec = torch.empty((B, I, REC), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        rc = lib.la3d_fit_boxes_bits(_ptr(depth), _ptr(bits), _ptr(cc), _ptr(K), _ptr(ground), B, I, H, W, _method_id(method),
                                     int(yaw_steps), int(seed) & 0xFFFFFFFF, int(image_offset) & 0xFFFFFFFF, _ptr(ws),
                                     ws_bytes, _ptr(rec), int(out_dtype == torch.float64), _stream())
    _lib.check(rc, "la3d_fit_boxes_bits")
    return rec


def fit_boxes_all(depth, K, masks, ground=None, out_dtype=torch.float64, method="pca", yaw_steps=0, sink=None,
                  workspace=None):
    """Boxes from EVERY masked pixel (``la3d_fit_boxes_all_to``): the reference's ``estimate_bbox`` with its random
    500-point draw (``src/util_3dbox.py:123-125``) replaced by the identity, for ``method`` pca / convex_hull /
    sweep.  Deterministic, no generator involved; same record layout and status codes as :func:`fit_boxes`.
    Two launches: the mask scan (with the camera / ground preparation riding in its grid) and one CTA per box
    that reduces the footprint's moments, hull candidates and extents over all its pixels."""
    lib = _lib.load()
    depth = _need_cuda("depth", depth, torch.float32)
    K = _need_cuda("K", K, torch.float64)
    masks = _need_cuda("masks", masks)
    B, I, H, W = masks.shape
    if tuple(depth.shape) != (B, H, W) or tuple(K.shape) != (B, 3, 3):
        raise ValueError("depth / K do not match the mask stack")
    if ground is not None:
        ground = _need_cuda("ground", ground, torch.float64)
        if tuple(ground.shape) != (B, I, 3):
            raise ValueError(f"ground must be [{B},{I},3]")
    if out_dtype not in (torch.float32, torch.float64):
        raise TypeError("out_dtype must be float32 or float64")
    m8, is01 = _masks_u8(masks)
    dev = depth.device
    ws_bytes = int(lib.la3d_fit_workspace_bytes(B, I, H, W))
    ws = workspace if workspace is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if ws.numel() < ws_bytes:
        raise ValueError("workspace too small")
    rec = None
    if sink is None:
        rec = torch.empty((B, I, REC), dtype=out_dtype, device=dev)
        sink = _lib.make_sink([rec.data_ptr()], out_dtype == torch.float64)
    with torch.cuda.device(dev):
        rc = lib.la3d_fit_boxes_all_to(_ptr(depth), _ptr(m8), _ptr(K), _ptr(ground), B, I, H, W, is01, _method_id(method),
                                       int(yaw_steps), _ptr(ws), ws_bytes, ctypes.byref(sink), _stream())
    _lib.check(rc, "la3d_fit_boxes_all_to")
    return rec


def fit_all_points(depth, prep, bits, I, out_dtype=torch.float64, method="pca", yaw_steps=0):
    """The dense fit alone (``la3d_fit_all_points_to``) on bit planes that already exist
    (:func:`mask_scan` / :func:`rle_decode`) and the ``prep`` buffer of :func:`fit_prepare`."""
    lib = _lib.load()
    depth = _need_cuda("depth", depth, torch.float32)
    bits = _need_cuda("bits", bits, torch.int32)
    B, H, W = depth.shape
    if tuple(bits.shape) != (B * I, scan_layout(H, W)[1]):
        raise ValueError("bit planes do not match depth and I")
    rec = torch.empty((B, I, REC), dtype=out_dtype, device=depth.device)
    with torch.cuda.device(depth.device):
        sink = _lib.make_sink([rec.data_ptr()], out_dtype == torch.float64)
        rc = lib.la3d_fit_all_points_to(_ptr(depth), _ptr(prep), _ptr(bits), B, I, H, W, _method_id(method), int(yaw_steps),
                                        ctypes.byref(sink), _stream())
    _lib.check(rc, "la3d_fit_all_points_to")
    return rec


class RleBoxFitter:
    """``BoxFitter`` for masks given as COCO run-length annotations (``la3d_fit_boxes_rle``): one call =
    three launches (decode with the preparation riding in its grid, subsample ranks, fit); the byte masks
    never exist.  ``total_runs`` / ``max_runs``: capacity of the run arrays the plan will be called with."""

    def __init__(self, B, I, H, W, total_runs, max_runs, device="cuda", out_dtype=torch.float64):
        self.lib = _lib.load()
        self.shape = (int(B), int(I), int(H), int(W))
        self.device = torch.device(device)
        if out_dtype not in (torch.float32, torch.float64):
            raise TypeError("out_dtype must be float32 or float64")
        self.out_dtype = out_dtype
        self.max_runs = int(max_runs)
        self.ws_bytes = int(self.lib.la3d_fit_workspace_bytes(*self.shape))
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.ends_ws = torch.empty((max(int(total_runs), 1),), dtype=torch.int32, device=self.device)
        self.rle_status = torch.zeros((B * I,), dtype=torch.int32, device=self.device)
        self.records = torch.empty((B, I, REC), dtype=out_dtype, device=self.device)

    def __call__(self, depth, K, run_counts, run_offsets, ground=None, method="pca", yaw_steps=0, seed=0, image_offset=0,
                 out=None, sink=None):
        """Returns ``records[B,I,64]``; ``self.rle_status[B*I]`` holds the decoder's per-plane status.
        ``sink``: as for :class:`BoxFitter` (``la3d_fit_boxes_rle_to``)."""
        B, I, H, W = self.shape
        depth = _need_cuda("depth", depth, torch.float32, pinned_ok=True)
        K = _need_cuda("K", K, torch.float64)
        run_offsets = _need_cuda("run_offsets", run_offsets, torch.int64)
        run_counts = _need_cuda("run_counts", run_counts)
        if run_counts.dtype not in (torch.int32, torch.uint32):
            raise TypeError("run_counts must be int32 / uint32")
        if tuple(depth.shape) != (B, H, W) or tuple(K.shape) != (B, 3, 3) or run_offsets.numel() != B * I + 1:
            raise ValueError(f"shapes do not match the plan {self.shape}")
        if run_counts.numel() > self.ends_ws.numel():
            raise ValueError("more runs than the plan was sized for (total_runs)")
        if ground is not None:
            ground = _need_cuda("ground", ground, torch.float64)
            if tuple(ground.shape) != (B, I, 3):
                raise ValueError(f"ground must be [{B},{I},3]")
        rec = self.records if out is None else _need_cuda("out", out, self.out_dtype)
        counts_ptr = _ptr(run_counts) if run_counts.numel() else _ptr(self.ends_ws)
        if sink is None:
            sink = _lib.make_sink([rec.data_ptr()], self.out_dtype == torch.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.la3d_fit_boxes_rle_to(_ptr(depth), counts_ptr, _ptr(run_offsets), self.max_runs, _ptr(self.ends_ws),
                                                _ptr(K), _ptr(ground), B, I, H, W, _method_id(method), int(yaw_steps),
                                                int(seed) & 0xFFFFFFFF, int(image_offset) & 0xFFFFFFFF,
                                                _ptr(self.workspace), self.ws_bytes, _ptr(self.rle_status),
                                                ctypes.byref(sink), _stream())
        _lib.check(rc, "la3d_fit_boxes_rle_to")
        return rec


def fit_points(points, offsets, sample_idx=None, K=None, ground=None, method="pca", yaw_steps=0,
               out_dtype=torch.float64):
    """Boxes from explicit point sets (the reference's own way of calling ``estimate_bbox``).

    ``points[total,3]`` float64, ``offsets[n+1]`` int64, ``sample_idx[n,500]`` int32 (rows to
    keep for sets with more than 500 points), ``K[n,3,3]`` / ``ground[n,3]`` float64 or None.
    """
    lib = _lib.load()
    points = _need_cuda("points", points, torch.float64)
    offsets = _need_cuda("offsets", offsets, torch.int64)
    n = offsets.numel() - 1
    if sample_idx is not None:
        sample_idx = _need_cuda("sample_idx", sample_idx, torch.int32)
        assert tuple(sample_idx.shape) == (n, SUBSAMPLE)
    if K is not None:
        K = _need_cuda("K", K, torch.float64)
        assert tuple(K.shape) == (n, 3, 3)
    if ground is not None:
        ground = _need_cuda("ground", ground, torch.float64)
        assert tuple(ground.shape) == (n, 3)
    rec = torch.empty((n, REC), dtype=out_dtype, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.la3d_fit_points(_ptr(points), _ptr(offsets), _ptr(sample_idx), _ptr(K), _ptr(ground), n,
                                 _method_id(method), int(yaw_steps), _ptr(rec), int(out_dtype == torch.float64),
                                 _stream())
    _lib.check(rc, "la3d_fit_points")
    return rec


def project_points(points, K, k_index=None):
    """``points[n,3]`` float64, ``K[m,3,3]`` (or ``[3,3]``) -> ``uv[n,2]``: the reference's ``project_to_2d``
    (``src/util.py:227-229``) for a batch; ``k_index[n]`` int32 picks the intrinsics of each point."""
    lib = _lib.load()
    points = _need_cuda("points", points, torch.float64)
    K = _need_cuda("K", K, torch.float64)
    n = points.shape[0]
    if k_index is not None:
        k_index = _need_cuda("k_index", k_index, torch.int32)
    uv = torch.empty((n, 2), dtype=torch.float64, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.la3d_project_points(_ptr(points), _ptr(K), _ptr(k_index), n, _ptr(uv), _stream())
    _lib.check(rc, "la3d_project_points")
    return uv


def iou_matrix(boxes0, boxes1, off0=None, off1=None):
    """Pairwise ``iou2D`` (``src/tools/combine_results.py:111-123`` of the reference) of two xyxy box
    lists, float64, bit-identical to the Python floats.  Without offsets: ``[n0,4] x [n1,4] -> [n0,n1]``.
    With ``off0`` / ``off1`` (``[G+1]`` int64 row offsets): one block per scene, all scenes in one
    launch; returns ``(flat, out_off)`` with block ``g`` = ``flat[out_off[g]:out_off[g+1]].view(n0_g, n1_g)``."""
    lib = _lib.load()
    boxes0 = _need_cuda("boxes0", boxes0, torch.float64)
    boxes1 = _need_cuda("boxes1", boxes1, torch.float64)
    dev = boxes0.device
    single = off0 is None
    if single:
        off0 = torch.tensor([0, boxes0.shape[0]], dtype=torch.int64, device=dev)
        off1 = torch.tensor([0, boxes1.shape[0]], dtype=torch.int64, device=dev)
    off0 = _need_cuda("off0", off0, torch.int64)
    off1 = _need_cuda("off1", off1, torch.int64)
    G = off0.numel() - 1
    sizes = (off0[1:] - off0[:-1]) * (off1[1:] - off1[:-1])
    out_off = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    out_off[1:] = torch.cumsum(sizes, 0)
    total = int(out_off[-1].item())
    flat = torch.empty((max(total, 1),), dtype=torch.float64, device=dev)
    if total:
        with torch.cuda.device(dev):
            rc = lib.la3d_iou_matrix(_ptr(boxes0), _ptr(off0), _ptr(boxes1), _ptr(off1), _ptr(out_off), G, _ptr(flat),
                                     _stream())
        _lib.check(rc, "la3d_iou_matrix")
    flat = flat[:total]
    if single:
        return flat.view(boxes0.shape[0], boxes1.shape[0])
    return flat, out_off


def box2d_from_corners(corners, K, wh, k_index=None):
    """``bbox2D_proj`` / ``bbox2D_trunc`` (``src/tools/combine_results.py:234-252``) of boxes given by
    their corners ``[n,8,3]``; ``K[m,3,3]``, ``wh[m,2] = (W, H)``, ``k_index[n]`` int32 picks the image."""
    lib = _lib.load()
    corners = _need_cuda("corners", corners, torch.float64)
    K = _need_cuda("K", K, torch.float64)
    wh = _need_cuda("wh", wh, torch.float64)
    n = corners.shape[0]
    if k_index is not None:
        k_index = _need_cuda("k_index", k_index, torch.int32)
    proj = torch.empty((n, 4), dtype=torch.float64, device=corners.device)
    trunc = torch.empty((n, 4), dtype=torch.float64, device=corners.device)
    if n:
        with torch.cuda.device(corners.device):
            rc = lib.la3d_box2d_from_corners(_ptr(corners), _ptr(K), _ptr(k_index), _ptr(wh), n, _ptr(proj), _ptr(trunc),
                                             _stream())
        _lib.check(rc, "la3d_box2d_from_corners")
    return proj, trunc


def depth_scale_median(depth_map, depth_render, mask_bits, render_bits, H, W):
    """Per plane the median of ``depth_map / depth_render`` over ``mask & render_mask``
    (``align_to_depth_match``, ``src/util.py:473-486`` of the reference), float32, bit-exact with
    ``np.median``.  ``depth_map[B,H,W]``, ``depth_render[B,I,H,W]`` float32; bit planes ``[B*I, words]``
    from :func:`mask_scan`.  Returns ``(n_overlap[B,I] int32, scale[B,I] float32)``; ``scale`` is NaN
    where the overlap is empty (the reference returns the identity transform there)."""
    lib = _lib.load()
    depth_map = _need_cuda("depth_map", depth_map, torch.float32)
    depth_render = _need_cuda("depth_render", depth_render, torch.float32)
    mask_bits = _need_cuda("mask_bits", mask_bits, torch.int32)
    render_bits = _need_cuda("render_bits", render_bits, torch.int32)
    B, I = depth_render.shape[:2]
    planes = B * I
    if depth_map.shape[0] != B or mask_bits.shape[0] != planes or render_bits.shape[0] != planes:
        raise ValueError("shape mismatch between depth maps and bit planes")
    n = torch.empty((B, I), dtype=torch.int32, device=depth_map.device)
    scale = torch.empty((B, I), dtype=torch.float32, device=depth_map.device)
    with torch.cuda.device(depth_map.device):
        rc = lib.la3d_masked_ratio_median(_ptr(depth_map), _ptr(depth_render), _ptr(mask_bits), _ptr(render_bits), planes,
                                          I, H, W, _ptr(n), _ptr(scale), _stream())
    _lib.check(rc, "la3d_masked_ratio_median")
    return n, scale


def ransac_subset_fit(x, y, idx):
    """Least-squares slope through the origin of the pairs ``(x[idx], y[idx])`` (``la3d_ransac_subset_fit``): what
    ``LinearRegression(fit_intercept=False).fit`` computes for one RANSAC trial of the reference's ``align_depth``
    (``src/batch_scripts/depth.py:52-92``).  ``x``, ``y`` float32 CUDA vectors, ``idx`` int64 CUDA vector.  Returns
    the slope as a Python float rounded to float32 (one device synchronisation: the loop on the host needs it)."""
    lib = _lib.load()
    x = _need_cuda("x", x, torch.float32)
    y = _need_cuda("y", y, torch.float32)
    idx = _need_cuda("idx", idx, torch.int64)
    sums = torch.empty(2, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.la3d_ransac_subset_fit(_ptr(x), _ptr(y), _ptr(idx), idx.numel(), _ptr(sums), _stream())
    _lib.check(rc, "la3d_ransac_subset_fit")
    sxx, sxy = sums.tolist()
    return float(np.float32(sxy / sxx))


def ransac_classify(x, y, coef, threshold):
    """``stats[6]`` (NumPy float64) of one RANSAC trial over ALL pairs (``la3d_ransac_classify``): number of inliers
    (``|y - x*coef| <= threshold`` in float32), and over the inliers sum y, sum y^2, sum residual^2, sum x^2, sum x y."""
    lib = _lib.load()
    stats = torch.empty(6, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.la3d_ransac_classify(_ptr(x), _ptr(y), x.numel(), float(coef), float(threshold), _ptr(stats), _stream())
    _lib.check(rc, "la3d_ransac_classify")
    return stats.cpu().numpy()


def scale_fill(rel, mask, coef, fill=10000.0):
    """``out = where(mask, rel * coef, fill)`` in float32 (``la3d_scale_fill``); ``mask=None``: ``~isinf(rel)``."""
    lib = _lib.load()
    rel = _need_cuda("rel", rel, torch.float32)
    m8 = None
    if mask is not None:
        m8, _ = _masks_u8(_need_cuda("mask", mask))
    out = torch.empty_like(rel)
    with torch.cuda.device(rel.device):
        rc = lib.la3d_scale_fill(_ptr(rel), _ptr(m8), rel.numel(), float(coef), float(fill), _ptr(out), _stream())
    _lib.check(rc, "la3d_scale_fill")
    return out


def median_f32(values):
    """Exact ``np.median`` of a float32 CUDA vector (float32 result, even counts average the two middle values in
    float32): the radix-select kernel of the depth-scale row (``la3d_masked_ratio_median``) on ``values / 1``."""
    values = _need_cuda("values", values, torch.float32)
    n = values.numel()
    ones = torch.ones((1, 1, 1, n), dtype=torch.float32, device=values.device)
    bits, _ = mask_scan(torch.ones((1, 1, n), dtype=torch.bool, device=values.device))
    _, scale = depth_scale_median(values.view(1, 1, n), ones, bits, bits, 1, n)
    return scale[0, 0]


def to_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
