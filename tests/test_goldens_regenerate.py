"""Where the reference tree is present (the build container), every golden fixture under tests/golden/ must come
out of its generating script again, array for array: the fixtures are outputs of the UNMODIFIED reference
(imported in place by tests/live_reference.py), not hand-edited numbers.  Skipped on machines without
/root/reference (the GPU box): there the committed fixtures are what the parity tests read."""
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

import live_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCRIPTS = ["make_golden", "make_golden_align", "make_golden_dense", "make_golden_next", "make_golden_rle", "make_golden_ties"]

pytestmark = pytest.mark.skipif(not live_reference.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def regenerated(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("regen")
    shutil.copytree(os.path.join(ROOT, "tests"), tmp / "tests", ignore=shutil.ignore_patterns("__pycache__", "regress_*"))
    os.makedirs(tmp / "labelany3d_b200" / "dropin")            # two scripts also rewrite the category tables
    for f in os.listdir(tmp / "tests" / "golden"):
        if f.endswith((".npz", ".json")):
            os.remove(tmp / "tests" / "golden" / f)             # nothing stale can pass for regenerated
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for name in SCRIPTS:
        proc = subprocess.run([sys.executable, os.path.join("tests", "golden", name + ".py")], cwd=tmp, env=env,
                              capture_output=True, text=True, timeout=600)
        assert proc.returncode == 0, name + "\n" + proc.stdout[-2000:] + proc.stderr[-2000:]
    return str(tmp)


def _same(a, b):
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    return np.array_equal(a, b, equal_nan=True) if a.dtype.kind in "fc" else np.array_equal(a, b)


@pytest.mark.parametrize("name", ["golden_v1.npz", "golden_align_v1.npz", "golden_dense_v1.npz", "golden_next_v1.npz",
                                  "golden_rle_v1.npz", "golden_ties_v1.npz"])
def test_array_fixtures_regenerate_bit_for_bit(regenerated, name):
    with np.load(os.path.join(GOLDEN, name)) as have, np.load(os.path.join(regenerated, "tests", "golden", name)) as got:
        assert sorted(have.files) == sorted(got.files)
        differ = [k for k in have.files if not _same(have[k], got[k])]
        assert not differ, differ[:10]


def test_json_fixtures_and_category_tables_regenerate(regenerated):
    def load(base, *parts):
        with open(os.path.join(base, *parts)) as f:
            return json.load(f)
    strip = lambda lines: [re.sub(r"/\S*/", "<tmp>/", ln) for ln in lines]  # noqa: E731  (scratch directories)
    have, got = load(GOLDEN, "golden_combine_v1.json"), load(regenerated, "tests", "golden", "golden_combine_v1.json")
    assert have["output"] == got["output"] and strip(have["log"]) == strip(got["log"])
    assert load(GOLDEN, "golden_rle_loader_v1.json") == load(regenerated, "tests", "golden", "golden_rle_loader_v1.json")
    for table in ("coco_omni3d_categories.json", "coco_category_names.json"):
        assert load(ROOT, "labelany3d_b200", "dropin", table) == load(regenerated, "labelany3d_b200", "dropin", table)
