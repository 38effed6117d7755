"""CPU oracle for the annotation -> mask-stack row (input side of the path).  TEST INFRASTRUCTURE ONLY.

NumPy / pure-Python restatement of how the reference turns COCO / COCONUT annotations into the
``bool [I,H,W]`` mask stack the box-fitting path consumes: ``read_bounding_boxes_segmentations``
(``src/util.py:337-382``) with its two segmentation formats, COCO run-length encoding
(``mask_utils.decode``, ``:361-370``) and polygons (``create_boolean_mask_from_polygon``,
``:386-415``).  Like the other oracle modules this is a checker: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

Parity status
  * uncompressed RLE (a ``counts`` list, column-major runs starting with a 0-run): PINNED against the
    reference's own encoder ``binary_mask_to_rle`` (``src/download_coconut.py:167-175``), imported
    live by ``tests/golden/make_golden_rle.py``: ``rle_encode`` equals it and ``rle_decode`` inverts it.
  * ``read_bounding_boxes_segmentations``: PINNED against the unmodified reference function run with
    this module's codec standing in for the absent ``pycocotools`` (same script).
  * the compressed ``counts`` STRING codec (``rle_from_string`` / ``rle_to_string``) lives in
    pycocotools (pinned ``pycocotools==2.0``, ``docs/INSTALL.md:32``; absent from the reference tree and
    from this image): restated from its published algorithm (``common/maskApi.c``: ``rleFrString``,
    ``rleToString``, ``rleDecode``, ``rleFrBbox`` are the only entry points the call sites reach).
    PARITY UNPINNED for the string form: no golden vector exists in the reference; the tests hold
    hand-derived strings and the encode/decode round trip.

=================================  ====================================================
oracle function                    follows
=================================  ====================================================
``rle_encode``                     ``src/download_coconut.py:167-175``
``rle_decode``                     pycocotools ``rleDecode`` (column-major fill), call
                                   site ``src/util.py:367``
``rle_from_string`` / ``_to_``     pycocotools ``rleFrString`` / ``rleToString``
``decode_annotation_rle``          ``src/util.py:364-370``
``polygon_mask``                   ``src/util.py:386-415`` (OpenCV ``fillPoly`` itself,
                                   like the reference)
``read_bounding_boxes_segmentations``  ``src/util.py:337-382``
``rle_to_bits``                    the bit planes / quarter counts of ``la3d_mask_scan``
                                   for a run-length encoded stack
=================================  ====================================================
"""

from __future__ import annotations

from itertools import groupby

import numpy as np


class InvalidRLE(ValueError):
    """The runs cover more than ``h*w`` pixels.  pycocotools 2.0 writes past the buffer here
    (later releases raise "Invalid RLE mask representation"); this path refuses."""


# ------------------------------------------------------------------ run-length codec
def rle_encode(binary_mask):
    """Uncompressed COCO RLE of a ``[h,w]`` mask (``src/download_coconut.py:167-175``): runs over the
    column-major flattening, the first run counts zeros (so it is 0 when pixel (0,0) is set)."""
    binary_mask = np.asarray(binary_mask)
    counts = []
    for i, (value, elements) in enumerate(groupby(binary_mask.ravel(order="F"))):
        if i == 0 and value == 1:
            counts.append(0)
        counts.append(len(list(elements)))
    return {"counts": counts, "size": list(binary_mask.shape)}


def rle_encode_fast(binary_mask):
    """Same counts as :func:`rle_encode`, vectorised (used to build large synthetic inputs)."""
    flat = np.asarray(binary_mask).ravel(order="F") != 0
    if flat.size == 0:
        return {"counts": [], "size": list(np.asarray(binary_mask).shape)}
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    edges = np.concatenate(([0], change, [flat.size]))
    counts = np.diff(edges).tolist()
    if flat[0]:
        counts = [0] + counts
    return {"counts": counts, "size": list(np.asarray(binary_mask).shape)}


def rle_decode(counts, h, w):
    """``[h,w]`` uint8 mask from column-major runs (pycocotools ``rleDecode``): runs alternate 0, 1,
    0, ...; pixels after the last run stay 0."""
    counts = [int(c) for c in counts]
    if any(c < 0 for c in counts):
        raise InvalidRLE("negative run length")
    if sum(counts) > h * w:
        raise InvalidRLE("Invalid RLE mask representation")
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape((h, w), order="F")


def rle_from_string(s):
    """Counts from pycocotools' compressed string (``rleFrString``): 6-bit groups in ASCII 48..111,
    5 payload bits each (least significant group first), bit 0x20 = "more groups follow", sign
    extension from bit 0x10 of the last group; from the fourth count on the value is a difference
    to the count two places back."""
    if isinstance(s, str):
        s = s.encode("utf-8")
    cnts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x & 0xFFFFFFFF)             # stored as uint
    return cnts


def rle_to_string(counts):
    """Inverse of :func:`rle_from_string` (pycocotools ``rleToString``)."""
    out = bytearray()
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c5 = x & 0x1F
            x >>= 5
            more = (x != -1) if (c5 & 0x10) else (x != 0)
            if more:
                c5 |= 0x20
            out.append(c5 + 48)
    return bytes(out)


def decode_annotation_rle(seg):
    """``mask_utils.decode(rle).astype(bool)`` of ``src/util.py:364-367`` for one annotation dict
    (``counts`` a list, a str or bytes; ``size = [h, w]``)."""
    h, w = seg["size"]
    counts = seg["counts"]
    if isinstance(counts, (str, bytes)):
        counts = rle_from_string(counts)
    return rle_decode(counts, h, w).astype(bool)


# ------------------------------------------------------------------ polygons
def polygon_mask(image_size, segmentation):
    """``create_boolean_mask_from_polygon`` (``src/util.py:386-415``), polygon branch: vertices
    truncated to int32, OpenCV ``fillPoly`` with colour 1; returns ``(mask, get_maximum_height)``.
    ``image_size`` is what the caller passes, ``(width, height)``."""
    import cv2
    mask = np.zeros((image_size[1], image_size[0]), dtype=np.uint8)
    for polygon in segmentation:
        points = np.array(polygon).reshape(-1, 2).astype(np.int32)
        cv2.fillPoly(mask, [points], color=1)
    boolean_mask = mask.astype(bool)
    rows = np.where(np.any(boolean_mask, axis=1))[0]
    return boolean_mask, (0 if rows.size == 0 else rows[-1] - rows[0] + 1)


# ------------------------------------------------------------------ the loader
def analyze_mask(mask, boundary_threshold=10, scale_threshold=100):
    b = boundary_threshold
    total = np.sum(mask[:b, :]) + np.sum(mask[-b:, :]) + np.sum(mask[:, :b]) + np.sum(mask[:, -b:])
    return total >= 10, np.sum(mask) >= scale_threshold


def read_bounding_boxes_segmentations(annotations, image_size, category_names):
    """``src/util.py:337-382`` for a list of annotation dicts: crowd annotations are skipped, RLE
    masks use the number of non-empty rows as their height, polygon masks the first-to-last row
    extent; a mask is kept when ``height / image_height > 0.0625`` and it neither touches the
    10-pixel border band with 10 or more pixels nor is smaller than 100 pixels.  Returns
    ``(bboxes, masks[I,H,W] bool, arange(I), category names)``."""
    bboxes, masks, cats = [], [], []
    for annotation in annotations:
        if annotation["iscrowd"]:
            continue
        if "segmentation" in annotation:
            seg = annotation["segmentation"]
            if isinstance(seg, dict) and "counts" in seg:
                mask = decode_annotation_rle(seg)
                height = np.sum(np.any(mask, axis=1))
            else:
                mask, height = polygon_mask(image_size, seg)
            is_truncated, is_scaleable = analyze_mask(mask)
            if height / image_size[1] > 0.0625 and not is_truncated and is_scaleable:
                masks.append(mask)
                cats.append(annotation["category_id"])
                bboxes.append(annotation["bbox"])
    names = [category_names.get(c, "unknown") for c in cats]
    return bboxes, np.array(masks), np.arange(len(masks)), names


# ------------------------------------------------------------------ device layout of a decoded stack
def rle_to_bits(counts_list, h, w, chunk_px=512):
    """What ``la3d_rle_decode`` must produce for a list of run lists: ``(bits[P, chunks*16] uint32,
    chunk_counts[P, chunks] uint32, status[P] int32)`` in the layout of ``la3d_mask_scan`` (bit ``k`` of
    word ``j`` = row-major pixel ``32 j + k``; one byte per 128-pixel quarter of a 512-pixel chunk).
    ``status`` is 1 where the runs overflow the image (bits then hold the in-image part)."""
    hw = h * w
    chunks = (hw + chunk_px - 1) // chunk_px
    P = len(counts_list)
    bits = np.zeros((P, chunks * 16), dtype=np.uint32)
    cc = np.zeros((P, chunks), dtype=np.uint32)
    status = np.zeros(P, dtype=np.int32)
    for p, counts in enumerate(counts_list):
        flat = np.zeros(hw, dtype=np.uint8)
        pos, v = 0, 0
        for c in counts:
            c = int(c)
            if v:
                flat[min(pos, hw):min(pos + c, hw)] = 1
            pos += c
            v ^= 1
        status[p] = int(pos > hw)
        row_major = flat.reshape((h, w), order="F").ravel()
        padded = np.zeros(chunks * chunk_px, dtype=np.uint8)
        padded[:hw] = row_major
        bits[p] = np.packbits(padded, bitorder="little").view(np.uint32)
        q = padded.reshape(chunks, 4, chunk_px // 4).sum(axis=2).astype(np.uint32)
        cc[p] = q[:, 0] | (q[:, 1] << 8) | (q[:, 2] << 16) | (q[:, 3] << 24)
    return bits, cc, status
