"""RANSAC depth alignment (row f3, second half): the oracle against the records of the unmodified reference
(tests/golden/golden_align_v1.npz, tests/golden/make_golden_align.py).  Same seed -> the same draws from the
process-global generator (its state after the call is compared by digest), the same prints, the same error; slope and
map within 2e-6 relative (the reference's slope is LAPACK float32 least squares, the restatement's is
float32(sum xy / sum xx): parity is statistical by construction, see the oracle's header)."""
import contextlib
import hashlib
import io
import os

import numpy as np
import pytest

import align_cases
from oracle import la3d_oracle_align as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_align_v1.npz")) as z:
        return {k: z[k] for k in z.files}


def state_digest():
    st = np.random.get_state()
    return hashlib.sha256(st[1].tobytes() + str(st[2:]).encode()).hexdigest()


def check_against_golden(fn, gold, exact_state=True):
    for name, (rel, metric, mask, seed) in align_cases.cases().items():
        np.random.seed(seed)
        m = None if mask is None else mask.copy()
        if f"{name}/raises" in gold:
            with pytest.raises(ValueError) as exc, contextlib.redirect_stdout(io.StringIO()):
                fn(rel.copy(), metric.copy(), mask=m)
            assert str(exc.value) == str(gold[f"{name}/raises"])
        else:
            with contextlib.redirect_stdout(io.StringIO()) as out:
                got = fn(rel.copy(), metric.copy(), mask=m)
            assert out.getvalue() == str(gold[f"{name}/printed"]), name
            want = gold[f"{name}/out"]
            assert got.dtype == want.dtype and got.shape == want.shape
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(got), fin), name
            assert np.all(np.abs(got[fin] - want[fin]) <= 2e-6 * np.abs(want[fin]) + 1e-6), name
        if exact_state:
            assert state_digest() == str(gold[f"{name}/state"]), name     # same draws, same number of trials


def test_oracle_align_depth_matches_the_reference():
    check_against_golden(ora.align_depth, load())


def test_dynamic_max_trials_and_errors():
    assert ora.dynamic_max_trials(10, 10, 3, 0.99) == 0 or ora.dynamic_max_trials(10, 10, 3, 0.99) == 1
    assert ora.dynamic_max_trials(1, 1000, 200, 0.99) == float("inf")
    x = np.linspace(1, 2, 50, dtype=np.float32)
    with pytest.raises(ValueError):
        ora.ransac_slope(x, 2 * x, min_samples=60)
