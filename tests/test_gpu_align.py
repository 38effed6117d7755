"""GPU parity of the RANSAC depth alignment (dropin/depth_align.py over csrc/align.cu): against the oracle EXACTLY
(same slope in float32, same number of trials, same map: the kernels compute the restatement's arithmetic) and against
the unmodified reference's records within 2e-6 relative with the generator left in the reference's state."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

import align_cases
from oracle import la3d_oracle_align as ora
from test_oracle_align_golden import check_against_golden, load, state_digest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def depth_align():
    assert torch.cuda.is_available()
    import labelany3d_b200
    path = labelany3d_b200.dropin_path()
    if path not in sys.path:
        sys.path.insert(0, path)
    import depth_align as mod
    return mod


def test_align_depth_matches_the_reference_records(depth_align):
    check_against_golden(depth_align.align_depth, load())


def test_align_depth_equals_the_oracle_exactly(depth_align):
    for name, (rel, metric, mask, seed) in align_cases.cases().items():
        if name == "inf_under_mask":
            continue
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            want = ora.align_depth(rel.copy(), metric.copy(), mask=None if mask is None else mask.copy())
        s_want = state_digest()
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            got = depth_align.align_depth(rel.copy(), metric.copy(), mask=None if mask is None else mask.copy())
        assert state_digest() == s_want, name
        np.testing.assert_array_equal(got, want, err_msg=name)


def test_ransac_pieces(depth_align):
    from labelany3d_b200 import ops
    rng = np.random.RandomState(0)
    x = rng.uniform(1, 3, 5000).astype(np.float32)
    y = (1.7 * x + 0.01 * rng.standard_normal(5000)).astype(np.float32)
    y[::7] *= 2.5
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    assert float(ops.median_f32(yd)) == float(np.median(y))
    idx = rng.choice(5000, 1000, replace=False)
    assert ops.ransac_subset_fit(xd, yd, torch.as_tensor(idx).cuda()) == float(ora.slope(x[idx], y[idx]))
    thr = float(np.median(np.abs(y - np.median(y))))
    _, want = ora.classify(x, y, np.float32(1.7), np.float32(thr))
    got = ops.ransac_classify(xd, yd, 1.7, thr)
    assert got[0] == want[0]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    coef, info = depth_align.ransac_slope(xd, yd, random_state=3)
    cw, iw = ora.ransac_slope(x, y, random_state=3)
    assert coef == float(cw) and info["n_trials"] == iw["n_trials"] and info["n_inliers"] == iw["n_inliers"]
