"""Deterministic inputs for the annotation -> mask-stack tests (COCO run-length codec, polygon masks,
``read_bounding_boxes_segmentations``).  Shared by tests/golden/make_golden_rle.py (which feeds them
to the unmodified reference) and by the CPU / GPU test-suites."""
import numpy as np


def _ellipse(H, W, cy, cx, ry, rx, theta=0.0):
    v, u = np.mgrid[:H, :W]
    du, dv = u - cx, v - cy
    c, s = np.cos(theta), np.sin(theta)
    return ((du * c + dv * s) / rx) ** 2 + ((dv * c - du * s) / ry) ** 2 <= 1.0


def codec_masks():
    """List of (name, mask[H,W] bool): every shape class the decoder distinguishes (widths that are /
    are not multiples of 32, fewer than 32 columns, heights that are not multiples of 32, H*W not a
    multiple of 512) and the run structures that matter (first pixel set, last pixel set, maximal run
    count, one run)."""
    rng = np.random.RandomState(777)
    out = []
    for (H, W) in ((480, 640), (96, 128), (75, 101), (33, 31), (64, 32), (5, 7), (1, 1), (100, 1), (1, 100),
                   (40, 257)):
        tag = f"{H}x{W}"
        out.append((f"{tag}/empty", np.zeros((H, W), bool)))
        out.append((f"{tag}/full", np.ones((H, W), bool)))
        first = np.zeros((H, W), bool)
        first[0, 0] = True
        out.append((f"{tag}/first_pixel", first))
        last = np.zeros((H, W), bool)
        last[-1, -1] = True
        out.append((f"{tag}/last_pixel", last))
        out.append((f"{tag}/noise50", rng.rand(H, W) < 0.5))
        out.append((f"{tag}/noise2", rng.rand(H, W) < 0.02))
        if H >= 5 and W >= 5:
            out.append((f"{tag}/ellipse", _ellipse(H, W, H * 0.45, W * 0.55, H * 0.3, W * 0.25, 0.4)))
            chk = (np.add.outer(np.arange(H), np.arange(W)) & 1).astype(bool)
            out.append((f"{tag}/checker", chk))
            rows = np.zeros((H, W), bool)
            rows[H // 2] = True
            out.append((f"{tag}/one_row", rows))
            cols = np.zeros((H, W), bool)
            cols[:, W // 3] = True
            out.append((f"{tag}/one_column", cols))
    return out


# Hand-derived strings of the compressed form (oracle/la3d_oracle_rle.py:rle_from_string).  A count is
# stored in 5-bit groups, least significant first, each as chr(48 + group | 0x20 if more follow); from
# the fourth count on the DIFFERENCE to the count two places back is stored; the last group's bit 0x10
# is the sign.  Worked by hand:
#   16  -> groups 16 (bit 0x10 set, so a second group must follow to keep the sign clear), 0 -> "`0"
#   40  -> groups 8 (more), 1 -> "X1";     -35 (= 5 - 40) -> groups 29 (more), 30 (sign) -> "mN"
STRING_VECTORS = [
    ([], b""),
    ([5], b"5"),
    ([0, 3, 2], b"032"),
    ([15], b"?"),
    ([16], b"`0"),
    ([6, 1, 40, 4, 5, 4, 5, 4, 21], b"61X13mN000`0"),
]


def loader_scene():
    """One image worth of COCONUT-style annotations that exercises every branch of
    ``read_bounding_boxes_segmentations`` (``src/util.py:337-382``): crowd skip, compressed-string RLE,
    polygon with one and with two rings, too small, too short, border-touching, missing segmentation,
    unknown category.  Returns ``(annotations, image_size=(W,H), masks_for_rle)``; the RLE strings are
    filled in by the caller with the codec under test (``seg['counts']`` holds the uncompressed list and
    ``seg['_mask_key']`` the key of the mask it encodes)."""
    H, W = 240, 320
    masks = {
        "big": _ellipse(H, W, 120, 160, 60, 80, 0.3),
        "short": _ellipse(H, W, 100, 90, 6, 60),            # 13 rows of 240: height ratio below 0.0625
        "small": _ellipse(H, W, 60, 250, 4, 4),             # area below 100
        "border": _ellipse(H, W, 8, 160, 30, 40),           # touches the top band
        "crowd": _ellipse(H, W, 180, 60, 30, 30),
        "tall": _ellipse(H, W, 130, 260, 90, 20),
        "holes": _ellipse(H, W, 120, 100, 70, 50) & ~_ellipse(H, W, 120, 100, 30, 20),
    }
    annos = []

    def rle(key, cat, crowd=0):
        annos.append({"iscrowd": crowd, "category_id": cat, "bbox": [1.0, 2.0, 3.0 + len(annos), 4.0],
                      "segmentation": {"size": [H, W], "counts": None, "_mask_key": key}})

    rle("big", 1)
    rle("crowd", 3, crowd=1)
    rle("short", 17)
    rle("small", 44)
    rle("border", 62)
    annos.append({"iscrowd": 0, "category_id": 2, "bbox": [5.0, 5.0, 50.0, 60.0]})           # no segmentation key
    # polygons: float vertices are truncated to int32 by the reference
    annos.append({"iscrowd": 0, "category_id": 18, "bbox": [100.0, 40.0, 90.5, 150.25],
                  "segmentation": [[100.7, 40.2, 190.9, 60.0, 170.3, 190.8, 120.1, 170.5]]})
    annos.append({"iscrowd": 0, "category_id": 999, "bbox": [30.0, 30.0, 80.0, 120.0],       # unknown category
                  "segmentation": [[30.0, 30.0, 110.0, 35.0, 105.0, 150.0, 40.0, 140.0],
                                   [200.0, 100.0, 260.0, 100.0, 260.0, 200.0, 200.0, 200.0]]})
    annos.append({"iscrowd": 0, "category_id": 5, "bbox": [0.0, 0.0, 20.0, 20.0],             # polygon on the border
                  "segmentation": [[0.0, 0.0, 40.0, 0.0, 40.0, 40.0, 0.0, 40.0]]})
    annos.append({"iscrowd": 0, "category_id": 7, "bbox": [10.0, 100.0, 300.0, 10.0],         # polygon too short
                  "segmentation": [[20.0, 100.0, 300.0, 100.0, 300.0, 110.0, 20.0, 110.0]]})
    rle("tall", 64)
    rle("holes", 88)
    return annos, (W, H), masks
