"""Golden vectors for hull-edge ties, minted from the UNMODIFIED reference.  Run in the build container:

    python tests/golden/make_golden_ties.py

For every footprint of tests/tie_cases.py: the reference's `estimate_bbox(..., method='convex_hull')` outputs and
`_estimate_yaw_convex_hull`'s yaw (SciPy / Qhull vertex order).  Also stored: whether the closed-form oracle (hull
from the lexicographically smallest point, what the kernels do) picks the same yaw.  Where it does not, the
minimum edge area is attained by more than one edge within one unit in the last place - the reference's own choice
then depends on Qhull's starting vertex, an implementation detail this repository does not re-create; both boxes
have the same footprint area (asserted here and by the tests).
Output: tests/golden/golden_ties_v1.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import live_reference  # noqa: E402
import tie_cases  # noqa: E402
from oracle import la3d_oracle as orc  # noqa: E402


def main():
    util, box, comb = live_reference.load()
    G, same, differ = {}, 0, []
    for name, pc in tie_cases.cases().items():
        with live_reference.quiet():
            v, ctr, dim, Rc = box.estimate_bbox(pc.copy(), None, None, "convex_hull")
            yaw = box._estimate_yaw_convex_hull(pc.copy())
        G[f"{name}/pc"] = pc
        G[f"{name}/vertices"], G[f"{name}/center"], G[f"{name}/dims"], G[f"{name}/R_cam"] = v, ctr, np.array(dim), Rc
        G[f"{name}/yaw"] = np.array(yaw)
        # the library form of the oracle is the reference, bit for bit
        assert orc.yaw_from_hull(pc, impl="scipy") == yaw
        closed = orc.yaw_from_hull(pc, impl="closed")
        agree = abs(closed - yaw) < 1e-12
        G[f"{name}/closed_agrees"] = np.array(agree)
        if agree:
            same += 1
        else:
            areas = tie_cases.edge_areas(pc)
            tied = int((areas <= areas.min() * (1 + 4e-16)).sum())
            assert tied >= 2, (name, areas)                       # only a tie can make the two orders disagree
            d = orc.fit_details(pc, None, "convex_hull", impl="closed")
            ref_area, our_area = dim[0] * dim[2], d["dimension"][0] * d["dimension"][2]
            assert abs(ref_area - our_area) <= 1e-12 * ref_area, (name, ref_area, our_area)
            differ.append(name)
    G["names"] = np.array(sorted(tie_cases.cases()))
    out = os.path.join(HERE, "golden_ties_v1.npz")
    np.savez_compressed(out, **G)
    print(f"wrote {out}: {same} footprints resolved like the reference, {len(differ)} exact ties resolved differently: {differ}")


if __name__ == "__main__":
    main()
