"""Property tests (hypothesis) of the oracle restatements the CUDA kernels implement, against the
libraries the reference itself calls (scikit-learn PCA, SciPy ConvexHull, NumPy's legacy RandomState).
CPU only; bounded so the whole file runs in well under a minute."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import la3d_oracle as orc
from oracle import la3d_oracle_next as orn

FAST = settings(max_examples=60, deadline=None, derandomize=True)


@FAST
@given(seed=st.integers(0, 2**32 - 1), high=st.integers(501, 400000))
def test_mt19937_restatement_is_numpys_stream(seed, high):
    """init_genrand + tempering + masked rejection == np.random.RandomState(seed).randint(0, high, 500),
    and the stream position afterwards is NumPy's too (a second instance continues where the first stopped)."""
    gen = orc.LegacyMT19937(seed)
    rs = np.random.RandomState(seed)
    np.testing.assert_array_equal(orc.legacy_randint(gen, high, 500), rs.randint(0, high, 500))
    np.testing.assert_array_equal(orc.legacy_randint(gen, high // 2 + 2, 500), rs.randint(0, high // 2 + 2, 500))


@FAST
@given(seed=st.integers(0, 10**6), n=st.integers(2, 500), yaw=st.floats(-3.1, 3.1), tilt=st.floats(-0.4, 0.4),
       use_ground=st.booleans())
def test_closed_forms_match_the_libraries(seed, n, yaw, tilt, use_ground):
    """Closed-form PCA yaw (what fit.cu computes) vs scikit-learn; gift-wrapped hull search vs SciPy/Qhull."""
    rng = np.random.RandomState(seed)
    pc = rng.normal(size=(n, 3)) * np.array([1.0 + rng.rand(), 0.3, 0.2 + rng.rand()])
    c, s = np.cos(yaw), np.sin(yaw)
    pc = pc @ np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]).T + np.array([0.3, -0.1, 4.0])
    ground = np.array([tilt, -1.0, 0.5 * tilt]) if use_ground else None
    for method in ("pca", "convex_hull"):
        try:
            want = orc.estimate_bbox(pc, None, ground, method, impl="library")
        except ValueError as exc:                       # e.g. n == 1 after filtering: both must raise alike
            with pytest.raises(ValueError, match=str(exc)[:20]):
                orc.estimate_bbox(pc, None, ground, method, impl="closed")
            continue
        got = orc.estimate_bbox(pc, None, ground, method, impl="closed")
        for a, b in zip(got, want):
            a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
            # the footprint is anisotropic by construction (x scaled 1-2, z 0.2-1.2 before the yaw), so the
            # principal axis is well conditioned; fp16 corner rounding may flip one ulp (2e-3 at 4 m)
            np.testing.assert_allclose(a, b, rtol=0, atol=4e-3)


@FAST
@given(seed=st.integers(0, 10**6), h=st.integers(1, 40), w=st.integers(1, 70), b=st.integers(0, 45), p=st.floats(0.0, 1.0))
def test_mask_statistics_restatement(seed, h, w, b, p):
    rng = np.random.RandomState(seed)
    m = rng.rand(h, w) < p
    s = orn.mask_stats(m[None], b)[0]
    tr, sc = orn.analyze_mask(m, (w, h), 100, b)
    assert bool(s[1] + s[2] + s[3] + s[4] >= 10) == bool(tr) and bool(s[0] >= 100) == bool(sc)
    assert (s[6] - s[5] + 1 if s[7] else 0) == orn.get_maximum_height(m)
    assert s[7] == orn.rows_with_pixels(m)


@FAST
@given(seed=st.integers(0, 10**6), n=st.integers(1, 300))
def test_median_of_ratios_is_a_float32_order_statistic(seed, n):
    """What median.cu relies on: np.median of float32 data = the middle order statistic, or the float32
    mean (add, then halve) of the two middle ones."""
    rng = np.random.RandomState(seed)
    r = (rng.uniform(1.5, 6.0, n).astype(np.float32) / rng.uniform(0.5, 3.0, n).astype(np.float32))
    srt = np.sort(r)
    want = srt[n // 2] if n % 2 else np.float32(np.float32(srt[n // 2 - 1] + srt[n // 2]) / np.float32(2))
    got = np.median(r)
    assert got.dtype == np.float32 and got == want


@FAST
@given(seed=st.integers(0, 10**6), h=st.integers(1, 70), w=st.integers(1, 150), density=st.sampled_from([0.0, 0.02, 0.3, 0.5, 0.97, 1.0]))
def test_run_length_codec_round_trips(seed, h, w, density):
    """COCO run-length codec (oracle/la3d_oracle_rle.py) and the product's host-side helpers
    (labelany3d_b200/coco_rle.py): encode -> decode is the identity, the compressed string survives a round
    trip through both parsers, the run sum is h*w, and the device layout agrees with the decoded mask."""
    from labelany3d_b200 import coco_rle
    from oracle import la3d_oracle_rle as orr
    rng = np.random.RandomState(seed)
    mask = rng.rand(h, w) < density
    counts = orr.rle_encode_fast(mask)["counts"]
    assert counts == orr.rle_encode(mask.astype(np.uint8))["counts"] == coco_rle.runs_from_mask(mask).tolist()
    assert sum(counts) == h * w and all(c > 0 for c in counts[1:])
    assert np.array_equal(orr.rle_decode(counts, h, w).astype(bool), mask)
    text = orr.rle_to_string(counts)
    assert orr.rle_from_string(text) == counts == coco_rle.counts_from_string(text).tolist()
    bits, cc, status = orr.rle_to_bits([counts], h, w)
    flat = np.unpackbits(bits[0].view(np.uint8), bitorder="little").astype(bool)
    assert status[0] == 0 and np.array_equal(flat[:h * w].reshape(h, w), mask) and not flat[h * w:].any()
    assert int(cc[0].view(np.uint8).sum()) == int(mask.sum())
