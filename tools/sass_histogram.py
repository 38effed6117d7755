"""Opcode histogram per kernel of the built library (cuobjdump -sass; no GPU needed): the evidence that the
library is sm_100a code with bulk-copy (TMA) staging where the design says so.

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "labelany3d_b200", "lib", "libla3d_sm100a.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"la3d::\(anonymous namespace\)::|la3d::", "", kern)
        kern = re.sub(r"\(.*\)$", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS opcode histogram of `labelany3d_b200/lib/libla3d_sm100a.so` (static instruction counts)\n")
print(f"`cuobjdump -sass`; cubin architectures: {', '.join(sorted(arch))}.  TMA bulk copies show as `UBLKCP`, mbarrier "
      "operations as `SYNCS`, warp reductions as `REDUX`; there is no `HMMA` / `UTC*MMA` anywhere (no dense contraction "
      "on this path) and no `UTMALDG` (the tiles are 1-D: `cp.async.bulk`, not tensor maps).\n")
print("| kernel | instructions | UBLKCP | SYNCS | LDG | STG | LDS | STS | DFMA+DMUL+DADD | DSETP | SHFL | REDUX | ATOM* | top opcodes |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, h in hist.items():
    tot = sum(h.values())
    top = ", ".join(f"{o} {n}" for o, n in h.most_common(6))
    atom = sum(n for o, n in h.items() if o.startswith(("ATOM", "RED")) and o != "REDUX")
    print(f"| `{k[:90]}` | {tot} | {h['UBLKCP']} | {h['SYNCS']} | {h['LDG']} | {h['STG']} | {h['LDS']} | {h['STS']} | "
          f"{h['DFMA'] + h['DMUL'] + h['DADD']} | {h['DSETP']} | {h['SHFL']} | {h['REDUX']} | {atom} | {top} |")
