"""labelany3d_b200 - B200 (sm_100a) implementation of LabelAny3D's 3D box-fitting hot path.

``ops``      torch-facing batched API over the C ABI (``include/la3d.h``)
``dropin``   directory holding ``util.py`` / ``util_3dbox.py`` / ``cam_utils.py`` with the
             reference's module names and function signatures (put it on ``sys.path``)
``dist``     image sharding across the GPUs of one node + the final all-gather
``synth``    synthetic COCO-shape inputs for tests and benchmarks
"""

import os as _os

__version__ = "0.1.0"


def dropin_path():
    """Directory to put in front of ``sys.path`` so ``import util_3dbox`` resolves here."""
    return _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")
