"""CPU pins of oracle/la3d_oracle_rle.py against what the unmodified reference produced
(tests/golden/make_golden_rle.py; no GPU, no reference tree needed)."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

import rle_cases
from oracle import la3d_oracle_rle as orr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_rle_v1.npz")) as z:
        return {k: z[k] for k in z.files}


def _digest(counts):
    return hashlib.sha256(np.asarray(counts, dtype=np.int64).tobytes()).hexdigest()


def test_encoder_matches_the_reference_encoder_and_decoder_inverts_it(gold):
    for name, mask in rle_cases.codec_masks():
        counts = orr.rle_encode_fast(mask)["counts"]
        assert len(counts) == int(gold[f"codec/{name}/n_runs"]), name
        assert _digest(counts) == str(gold[f"codec/{name}/sha256"]), name
        if f"codec/{name}/counts" in gold:
            ref_counts = gold[f"codec/{name}/counts"].tolist()
            assert counts == ref_counts, name
            assert orr.rle_encode(mask.astype(np.uint8))["counts"] == ref_counts, name
        h, w = mask.shape
        assert np.array_equal(orr.rle_decode(counts, h, w).astype(bool), mask), name
        assert counts == [] or sum(counts) == h * w, name


def test_compressed_string_codec():
    for counts, text in rle_cases.STRING_VECTORS:
        assert orr.rle_to_string(counts) == text
        assert orr.rle_from_string(text) == counts
        assert orr.rle_from_string(text.decode("ascii")) == counts
    rng = np.random.RandomState(5)
    for _ in range(50):
        n = rng.randint(0, 40)
        counts = rng.randint(0, 2 ** rng.randint(1, 20), size=n).tolist()
        assert orr.rle_from_string(orr.rle_to_string(counts)) == counts
    for name, mask in rle_cases.codec_masks()[:40]:
        counts = orr.rle_encode_fast(mask)["counts"]
        assert orr.rle_from_string(orr.rle_to_string(counts)) == counts, name


def test_decoder_refuses_runs_beyond_the_image():
    with pytest.raises(orr.InvalidRLE):
        orr.rle_decode([0, 13], 3, 4)
    assert orr.rle_decode([0, 12], 3, 4).all()
    assert not orr.rle_decode([], 3, 4).any()
    short = orr.rle_decode([2, 3], 3, 4)                      # pixels after the last run stay 0
    assert short.ravel(order="F").tolist() == [0, 0, 1, 1, 1] + [0] * 7


def test_loader_matches_the_reference(gold):
    annos, size, masks = rle_cases.loader_scene()
    for a in annos:
        seg = a.get("segmentation")
        if isinstance(seg, dict):
            seg["counts"] = orr.rle_to_string(orr.rle_encode_fast(masks[seg.pop("_mask_key")])["counts"]).decode("ascii")
    with open(os.path.join(ROOT, "labelany3d_b200", "dropin", "coco_category_names.json")) as f:
        names = {int(k): v for k, v in json.load(f).items()}
    with open(os.path.join(ROOT, "tests", "golden", "golden_rle_loader_v1.json")) as f:
        want = json.load(f)
    before = copy.deepcopy(annos)
    bboxes, stack, ids, cats = orr.read_bounding_boxes_segmentations(annos, size, names)
    assert annos == before                                     # inputs are not modified
    assert bboxes == want["bboxes"] and cats == want["names"]
    shape = tuple(gold["loader/masks_shape"])
    ref_stack = np.unpackbits(gold["loader/masks_packed"], bitorder="little")[:int(np.prod(shape))].reshape(shape).astype(bool)
    assert stack.dtype == bool and np.array_equal(stack, ref_stack)
    np.testing.assert_array_equal(ids, gold["loader/ids"])


def test_device_layout_of_a_decoded_stack():
    """rle_to_bits (what la3d_rle_decode must write) = the scan layout of the decoded masks."""
    for (H, W) in ((75, 101), (64, 32), (5, 7), (40, 257)):
        cases = [m for n, m in rle_cases.codec_masks() if n.startswith(f"{H}x{W}/")]
        runs = [orr.rle_encode_fast(m)["counts"] for m in cases]
        bits, cc, status = orr.rle_to_bits(runs, H, W)
        assert not status.any()
        chunks = (H * W + 511) // 512
        for p, m in enumerate(cases):
            flat = np.zeros(chunks * 512, dtype=bool)
            flat[:H * W] = m.ravel()
            assert np.array_equal(np.unpackbits(bits[p].view(np.uint8), bitorder="little").astype(bool), flat)
            assert np.array_equal(cc[p].view(np.uint8).reshape(chunks, 4), flat.reshape(chunks, 4, 128).sum(-1))
    bits, cc, status = orr.rle_to_bits([[0, 40], [3, 2]], 5, 7)
    assert status.tolist() == [1, 0] and int(cc[0].view(np.uint8).sum()) == 35 and int(cc[1].view(np.uint8).sum()) == 2
