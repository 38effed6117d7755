"""Golden vectors for the RANSAC depth alignment, minted from the UNMODIFIED reference.  Run in the build container:

    python tests/golden/make_golden_align.py

`align_depth` lives in `src/batch_scripts/depth.py:52-92`, a script whose top level imports the depth networks; the
function itself needs only NumPy and scikit-learn.  This script therefore takes the function's SOURCE TEXT out of the
reference file in place (ast, nothing is copied into the repository) and executes it unmodified, with the process-global
NumPy generator seeded per case (the reference never seeds it; `RANSACRegressor(random_state=None)` draws from it).
Stored per case of tests/align_cases.py: the output map, the fitted slope (recovered as output / input under the mask)
and a digest of the generator's state after the call.  The oracle (`oracle/la3d_oracle_align.py`) is asserted against
it here: same generator state, slope within 2e-6 relative, map within 2e-6 relative.
Output: tests/golden/golden_align_v1.npz.
"""
import ast
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import align_cases  # noqa: E402
from oracle import la3d_oracle_align as ora  # noqa: E402

REF = "/root/reference/src/batch_scripts/depth.py"


def reference_align_depth():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "align_depth")
    ns = {}
    from sklearn.linear_model import LinearRegression, RANSACRegressor
    ns.update(np=np, RANSACRegressor=RANSACRegressor, LinearRegression=LinearRegression)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["align_depth"]


def state_digest():
    st = np.random.get_state()
    return hashlib.sha256(st[1].tobytes() + str(st[2:]).encode()).hexdigest()


def main():
    ref = reference_align_depth()
    G = {}
    for name, (rel, metric, mask, seed) in align_cases.cases().items():
        np.random.seed(seed)
        try:
            with contextlib.redirect_stdout(io.StringIO()) as out:
                got = ref(rel.copy(), metric.copy(), mask=None if mask is None else mask.copy())
        except ValueError as exc:
            G[f"{name}/raises"] = np.array(str(exc))
            G[f"{name}/state"] = np.array(state_digest())
            np.random.seed(seed)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    ora.align_depth(rel.copy(), metric.copy(), mask=None if mask is None else mask.copy())
                raise AssertionError(f"{name}: the oracle did not raise")
            except ValueError as mine:
                assert str(mine) == str(exc), (str(mine), str(exc))
            assert state_digest() == str(G[f"{name}/state"]), name
            print(f"{name}: raises {str(exc)!r}")
            continue
        G[f"{name}/out"] = got
        G[f"{name}/printed"] = np.array(out.getvalue())
        G[f"{name}/state"] = np.array(state_digest())
        sel = (mask if mask is not None else ~np.isinf(rel)) & np.isfinite(rel) & (rel != 0)
        if got is metric or np.array_equal(got, metric):
            G[f"{name}/coef"] = np.array(np.nan)
        else:
            G[f"{name}/coef"] = np.array(np.median(got[sel].astype(np.float64) / rel[sel].astype(np.float64)))
        # the oracle on the same seed
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()) as out2:
            mine, info = ora.align_depth(rel.copy(), metric.copy(), mask=None if mask is None else mask.copy(), return_info=True)
        assert state_digest() == str(G[f"{name}/state"]), name            # the same draws, the same number of them
        assert out2.getvalue() == out.getvalue(), (name, out2.getvalue(), out.getvalue())
        if info["coef"] is None:
            assert np.array_equal(mine, got), name
        else:
            rel_err = abs(info["coef"] - float(G[f"{name}/coef"])) / abs(float(G[f"{name}/coef"]))
            assert rel_err < 2e-6, (name, info["coef"], float(G[f"{name}/coef"]))
            fin = np.isfinite(got)
            assert np.array_equal(np.isfinite(mine), fin)
            assert np.all(np.abs(mine[fin] - got[fin]) <= 2e-6 * np.abs(got[fin]) + 1e-6), name
            print(f"{name}: reference slope {float(G[name + '/coef']):.8f}, oracle {info['coef']:.8f} (rel {rel_err:.1e}), "
                  f"{info['n_trials']} trials, {info['n_inliers']} / {info['n']} inliers")
    out_path = os.path.join(HERE, "golden_align_v1.npz")
    np.savez_compressed(out_path, **G)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
