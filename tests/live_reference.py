"""Import the UNMODIFIED reference functions in place (read-only) for CPU-side pinning.

Only usable where ``/root/reference`` exists (the build container).  The GPU box
does not have it, so nothing marked ``gpu`` may call :func:`load`.

The reference modules are loaded under private names (``_la3d_ref_util`` ...)
so they never collide with the drop-in modules ``util`` / ``util_3dbox`` that
this repository ships.  ``trimesh``, ``rembg`` and ``pycocotools`` are absent
from this image and are only needed by reference functions outside the hot
path, so empty stand-ins are put in ``sys.modules`` for the import (SURVEY.md
section 8c).
"""

from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "util_3dbox.py"))


_cache = {}


def load():
    """Returns ``(util, util_3dbox, combine_results)`` reference modules."""
    if _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not present")
    added = []
    for name in ("trimesh", "rembg", "pycocotools", "pycocotools.mask"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            added.append(name)
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]

    def _load(alias, rel):
        spec = importlib.util.spec_from_file_location(alias, os.path.join(REF_SRC, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    try:
        mods = (_load("_la3d_ref_util", "util.py"),
                _load("_la3d_ref_util_3dbox", "util_3dbox.py"),
                _load("_la3d_ref_combine", "tools/combine_results.py"))
    finally:
        for name in added:
            sys.modules.pop(name, None)
    _cache["mods"] = mods
    return mods


@contextlib.contextmanager
def quiet():
    """The reference prints one line per box; keep test logs readable."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
