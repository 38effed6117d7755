"""CPU checks of the host side of the run-length row: the compressed-string parser of
``labelany3d_b200/coco_rle.py`` (product code; NumPy, vectorised) against the oracle's restatement of
pycocotools' ``rleFrString`` and the hand-derived strings, and the flat run arrays the kernel takes."""
import numpy as np
import pytest

import rle_cases
from labelany3d_b200 import coco_rle
from oracle import la3d_oracle_rle as orr


def test_string_parser_matches_the_hand_derived_vectors_and_the_oracle():
    for counts, text in rle_cases.STRING_VECTORS:
        assert coco_rle.counts_from_string(text).tolist() == counts
        assert coco_rle.counts_from_string(text.decode("ascii")).tolist() == counts
    rng = np.random.RandomState(11)
    for _ in range(300):
        n = rng.randint(0, 40)
        counts = rng.randint(0, 2 ** rng.randint(1, 31), size=n).tolist()
        text = orr.rle_to_string(counts)
        assert coco_rle.counts_from_string(text).tolist() == counts == orr.rle_from_string(text)
    for name, mask in rle_cases.codec_masks():
        counts = orr.rle_encode_fast(mask)["counts"]
        assert coco_rle.counts_from_string(orr.rle_to_string(counts)).tolist() == counts, name
    # malformed strings: the C parser wraps negative results to unsigned, and so do both restatements
    for text in (b"5N", b"00N", b"0000N"):
        assert coco_rle.counts_from_string(text).tolist() == orr.rle_from_string(text)
    with pytest.raises(ValueError):
        coco_rle.counts_from_string(b"5P")                    # the last character announces another one


def test_runs_from_mask_is_the_reference_encoding():
    for name, mask in rle_cases.codec_masks():
        assert coco_rle.runs_from_mask(mask).tolist() == orr.rle_encode_fast(mask)["counts"], name
    assert coco_rle.runs_from_mask(np.zeros((0, 0), bool)).size == 0


def test_runs_of_and_pack_runs():
    runs, size = coco_rle.runs_of({"size": [3, 4], "counts": "246"})
    assert runs.dtype == np.uint32 and runs.tolist() == [2, 4, 6] and size == (3, 4)
    runs, size = coco_rle.runs_of({"size": [3, 4], "counts": [0, 12]})
    assert runs.tolist() == [0, 12]
    with pytest.raises(ValueError):
        coco_rle.runs_of({"size": [3, 4], "counts": [1, -2]})
    counts, offsets, max_runs = coco_rle.pack_runs([np.array([1, 2, 3], np.uint32), np.zeros(0, np.uint32), np.array([9], np.uint32)])
    assert counts.tolist() == [1, 2, 3, 9] and offsets.tolist() == [0, 3, 3, 4] and max_runs == 3
    counts, offsets, max_runs = coco_rle.pack_runs([])
    assert counts.size == 0 and offsets.tolist() == [0] and max_runs == 0
    counts, offsets, max_runs = coco_rle.pack_runs([np.zeros(0, np.uint32)])
    assert counts.size == 0 and offsets.tolist() == [0, 0] and max_runs == 0
