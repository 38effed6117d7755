"""Host side of the COCO run-length format: the compressed ``counts`` string of COCO / COCONUT
annotations -> run lengths, and run lists -> the flat arrays ``la3d_rle_decode`` takes.

The reference hands the string to pycocotools (``mask_utils.decode``, ``src/util.py:364-367``), whose
``rleFrString`` parses it on the host as well; the parsing below is that format in vectorised NumPy
(5 payload bits per character starting at ASCII 48, bit ``0x20`` = "more characters follow", bit
``0x10`` of the last character = sign, and from the fourth run on the stored value is the difference
to the run two places back).  Turning runs into masks is the GPU's job (``csrc/rle.cu``).
"""

from __future__ import annotations

import numpy as np


def counts_from_string(s):
    """Run lengths (``uint32`` array) of a compressed ``counts`` string (``str`` or ``bytes``)."""
    if isinstance(s, str):
        s = s.encode("utf-8")
    c = np.frombuffer(s, dtype=np.uint8).astype(np.int64) - 48
    if c.size == 0:
        return np.zeros(0, dtype=np.uint32)
    more = (c & 0x20) != 0
    if more[-1]:
        raise ValueError("truncated RLE string: the last character announces another one")
    last = ~more                                           # last character of every run
    run_id = np.concatenate(([0], np.cumsum(last)[:-1]))   # which run a character belongs to
    starts = np.flatnonzero(np.concatenate(([True], last[:-1])))
    k = np.arange(c.size) - starts[run_id]                 # position of the character inside its run
    if k.max() > 12:
        raise ValueError("RLE string holds a run of more than 13 characters (beyond 64 bits)")
    n = int(run_id[-1]) + 1
    x = np.zeros(n, dtype=np.int64)
    np.add.at(x, run_id, (c & 0x1F) << (5 * k))
    neg = last & ((c & 0x10) != 0)                          # sign bit of the last character: extend it
    x[run_id[neg]] |= np.int64(-1) << (5 * (k[neg] + 1))
    # runs 3, 5, 7, ... add to run 1's chain, runs 4, 6, ... to run 2's; run 0 stands alone
    out = x.copy()
    out[1::2] = np.cumsum(x[1::2])
    if n > 2:
        out[2::2] = np.cumsum(x[2::2])
    return (out & 0xFFFFFFFF).astype(np.uint32)


def runs_of(segmentation):
    """``(run lengths uint32, (h, w))`` of one annotation's RLE dict (``counts``: str, bytes or list)."""
    h, w = (int(v) for v in segmentation["size"])
    counts = segmentation["counts"]
    if isinstance(counts, (str, bytes)):
        runs = counts_from_string(counts)
    else:
        runs = np.asarray(counts, dtype=np.int64)
        if runs.size and (runs.min() < 0 or runs.max() > 0xFFFFFFFF):
            raise ValueError("run lengths must fit an unsigned 32-bit integer")
        runs = runs.astype(np.uint32)
    return runs, (h, w)


def runs_from_mask(mask):
    """Run lengths (``uint32``) of one ``[h,w]`` mask in COCO order (column-major, first run counts
    zeros): the format of the reference's ``binary_mask_to_rle`` (``src/download_coconut.py:167-175``)."""
    flat = np.asarray(mask).ravel(order="F") != 0
    if flat.size == 0:
        return np.zeros(0, dtype=np.uint32)
    edges = np.concatenate(([0], np.flatnonzero(flat[1:] != flat[:-1]) + 1, [flat.size]))
    runs = np.diff(edges)
    if flat[0]:
        runs = np.concatenate(([0], runs))
    return runs.astype(np.uint32)


def pack_runs(run_lists):
    """``(counts uint32[total], offsets int64[P+1], max_runs)`` for a list of run-length arrays."""
    sizes = np.array([len(r) for r in run_lists], dtype=np.int64)
    offsets = np.zeros(len(run_lists) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    counts = (np.concatenate([np.asarray(r, dtype=np.uint32) for r in run_lists]) if len(run_lists) and offsets[-1]
              else np.zeros(0, dtype=np.uint32))
    return counts, offsets, int(sizes.max()) if len(run_lists) else 0


def runs_from_masks_device(masks):
    """:func:`runs_from_mask` + :func:`pack_runs` for a whole stack on its device with torch:
    ``masks[P,H,W]`` (bool / uint8 tensor) -> ``(counts int32[total] tensor, offsets int64[P+1] tensor,
    max_runs int)``.  Input preparation for benchmarks and tests (the annotations of a real run come from
    COCO JSON); equal to the NumPy form, which a GPU test asserts."""
    import torch
    P, H, W = masks.shape
    HW = H * W
    flat = (masks != 0).transpose(1, 2).reshape(P, HW)              # column-major pixel order
    first = flat[:, 0].to(torch.int64)                               # a plane that starts set gets a leading 0-run
    change = flat[:, 1:] != flat[:, :-1]
    c = change.sum(dim=1, dtype=torch.int64)
    nruns = c + 1 + first
    offsets = torch.zeros(P + 1, dtype=torch.int64, device=masks.device)
    torch.cumsum(nruns, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    ends = torch.empty(total, dtype=torch.int64, device=masks.device)
    idx = change.nonzero()                                           # sorted by plane, then position
    plane, pos = idx[:, 0], idx[:, 1] + 1
    cstart = torch.cumsum(c, 0) - c                                  # exclusive prefix of the change counts
    k = torch.arange(idx.shape[0], device=masks.device) - cstart[plane]
    ends[offsets[:-1][plane] + first[plane] + k] = pos
    ends[offsets[1:] - 1] = HW
    lead = first.nonzero()[:, 0]
    ends[offsets[:-1][lead]] = 0
    prev = torch.cat([ends.new_zeros(1), ends[:-1]])
    prev[offsets[:-1]] = 0
    counts = (ends - prev).to(torch.int32)
    return counts, offsets, int(nruns.max().item())
