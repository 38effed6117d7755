"""Golden vectors for the annotation -> mask-stack row, minted from the UNMODIFIED reference.
Run in the build container:

    python tests/golden/make_golden_rle.py

* imports ``src/download_coconut.py`` of the reference in place (stand-ins only for its absent
  ``skimage`` / ``datasets`` / ``pycocotools`` imports) and runs its run-length ENCODER
  ``binary_mask_to_rle`` (``:167-175``) on the masks of ``tests/rle_cases.py``: the run lists (or, for
  the long ones, their SHA-256) are stored and the oracle codec (``oracle/la3d_oracle_rle.py``) is
  asserted to produce the same runs and to invert them;
* runs the reference's ``read_bounding_boxes_segmentations`` (``src/util.py:337-382``) unmodified on the
  annotation scene of ``rle_cases.loader_scene()``, with the oracle codec standing in for
  ``pycocotools.mask.decode`` (pycocotools is not in this image; its string format is restated from the
  published algorithm and is the one part of this row whose parity is unpinned), stores the outputs and
  asserts the oracle's restatement of the loader equal;
* writes the COCO / COCONUT category-name table the loader returns names from (an interface constant
  of the reference, ``src/util.py:419-449``) to ``labelany3d_b200/dropin/coco_category_names.json``.

Outputs: ``tests/golden/golden_rle_v1.npz``, ``tests/golden/golden_rle_loader_v1.json``.
"""
import contextlib
import copy
import hashlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import live_reference  # noqa: E402
import rle_cases  # noqa: E402
from oracle import la3d_oracle_rle as orr  # noqa: E402

FULL_COUNTS_BELOW = 4096          # run lists shorter than this are stored whole


def counts_digest(counts):
    return hashlib.sha256(np.asarray(counts, dtype=np.int64).tobytes()).hexdigest()


def load_reference_encoder():
    added = []
    for name in ("skimage", "skimage.measure", "datasets", "pycocotools", "pycocotools.mask"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            added.append(name)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["datasets"].load_dataset = None
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    try:
        spec = importlib.util.spec_from_file_location("_la3d_ref_download_coconut",
                                                      os.path.join(live_reference.REF_SRC, "download_coconut.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in added:
            sys.modules.pop(name, None)
    return mod


def standin_decode(rle):
    """``pycocotools.mask.decode`` for one dict with compressed ``counts`` bytes: a Fortran-ordered
    uint8 ``[h,w]`` array, like the original."""
    assert isinstance(rle["counts"], bytes), "pycocotools takes the compressed string as bytes"
    h, w = rle["size"]
    return np.asfortranarray(orr.rle_decode(orr.rle_from_string(rle["counts"]), h, w))


def scene_with_strings():
    annos, size, masks = rle_cases.loader_scene()
    for a in annos:
        seg = a.get("segmentation")
        if isinstance(seg, dict):
            counts = orr.rle_encode(masks[seg.pop("_mask_key")].astype(np.uint8))["counts"]
            seg["counts"] = orr.rle_to_string(counts).decode("ascii")       # COCONUT stores a str
    return annos, size


def main():
    G = {}
    ref_enc = load_reference_encoder()
    for name, mask in rle_cases.codec_masks():
        ref = ref_enc.binary_mask_to_rle(mask.astype(np.uint8))
        assert ref["size"] == list(mask.shape)
        counts = [int(c) for c in ref["counts"]]
        G[f"codec/{name}/n_runs"] = np.asarray(len(counts))
        G[f"codec/{name}/sha256"] = np.asarray(counts_digest(counts))
        if len(counts) < FULL_COUNTS_BELOW:
            G[f"codec/{name}/counts"] = np.asarray(counts, dtype=np.int64)
        # the oracle codec against the reference's encoder
        assert orr.rle_encode(mask.astype(np.uint8))["counts"] == counts, name
        assert orr.rle_encode_fast(mask)["counts"] == counts, name
        h, w = mask.shape
        assert np.array_equal(orr.rle_decode(counts, h, w).astype(bool), mask), name
        assert orr.rle_from_string(orr.rle_to_string(counts)) == counts, name
    for counts, text in rle_cases.STRING_VECTORS:
        assert orr.rle_to_string(counts) == text and orr.rle_from_string(text) == counts

    util, _, _ = live_reference.load()
    util.mask_utils.decode = standin_decode
    annos, size = scene_with_strings()
    with contextlib.redirect_stdout(io.StringIO()) as log:
        bboxes, masks, ids, names = util.read_bounding_boxes_segmentations(copy.deepcopy(annos), size)
    o_b, o_m, o_i, o_n = orr.read_bounding_boxes_segmentations(copy.deepcopy(annos), size, dict(util.COCO_CATEGORIES))
    assert o_b == bboxes and o_n == names and np.array_equal(o_i, ids)
    assert o_m.dtype == masks.dtype == bool and np.array_equal(o_m, masks)
    G["loader/masks_packed"] = np.packbits(masks, axis=None, bitorder="little")
    G["loader/masks_shape"] = np.asarray(masks.shape)
    G["loader/ids"] = np.asarray(ids)
    with open(os.path.join(HERE, "golden_rle_loader_v1.json"), "w") as f:
        json.dump({"bboxes": bboxes, "names": names, "log": log.getvalue().splitlines()}, f, indent=1)
    # an empty annotation list: np.array([]) and np.arange(0)
    b0, m0, i0, n0 = util.read_bounding_boxes_segmentations([], size)
    assert b0 == [] and n0 == [] and m0.shape == (0,) and i0.shape == (0,)

    np.savez_compressed(os.path.join(HERE, "golden_rle_v1.npz"), **G)
    with open(os.path.join(ROOT, "labelany3d_b200", "dropin", "coco_category_names.json"), "w") as f:
        json.dump({str(k): v for k, v in util.COCO_CATEGORIES.items()}, f, indent=0)
    print(f"{len(G)} arrays; loader kept {len(names)} of {len(annos)} annotations: {names}; reference printed {log.getvalue().splitlines()}")


if __name__ == "__main__":
    main()
