"""Aggregate an ncu source-page capture per CUDA source line (no GUI needed).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX MANGLED_SUBSTR [launch_skip] [top]

The ncu CSV source page is per SASS instruction without line numbers; nvdisasm -g on
the cubin extracted from the built .so gives the line of every instruction.  Both list
the function's instructions in address order, so they are zipped by index.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kregex, mangled = sys.argv[1:4]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("LA3D_SO", os.path.join(root, "labelany3d_b200", "lib", "libla3d_sm100a.so"))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kregex}", "-s", str(skip), "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address":
        break                      # ncu prints the listing twice
    if len(r) == len(hdr):
        sass.append(r)
print(rows[0][1] if rows and len(rows[0]) > 1 else "", f"-- {len(sass)} SASS instructions")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = []
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    if os.path.basename(cub).count("-") > 1:
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
    inside, cur, done = False, ("?", 0), False
    for ln in dis:
        if ln.startswith("//---") and ".text." in ln:
            if inside:
                done = True        # first matching function only
            inside = (mangled in ln) and not done
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(2).strip()))
    if lines:
        break
assert lines, "function not found in cubins"
if len(lines) != len(sass):
    print(f"warning: nvdisasm has {len(lines)} instructions, ncu has {len(sass)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for (loc, text), r in zip(lines, sass):
    a = agg[loc]
    a["samples"] += int(r[idx["# Samples"]] or 0)
    a["inst"] += int(r[idx["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += int(r[idx[s]] or 0)
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
src_cache = {}


def src(loc):
    f, n = loc
    for d in ("labelany3d_b200/csrc", "include"):
        p = os.path.join(os.environ.get("LA3D_SRC_ROOT", root), d, f)
        if os.path.isfile(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip() if 0 < n <= len(src_cache[p]) else ""
    return ""


print(f"total samples {tot}, warp instructions {toti}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = " ".join(f"{s[6:]}={a[s]}" for s in sorted(stalls, key=lambda s: -a[s])[:3] if a[s])
    print(f"{100 * a['samples'] / tot:5.1f}% smp {100 * a['inst'] / toti:5.1f}% inst  {loc[0]}:{loc[1]:<4d} {src(loc)[:70]:70s} | {st}")
