"""Markdown table of the metrics that matter from an ``ncu --set full`` report (no GPU needed).

    python tools/ncu_summary.py REPORT.ncu-rep [--traffic-json OUT.json] > profiles/NAME.md

One column per captured launch.  ``--traffic-json`` also writes ``{kernel: dram bytes per launch}``
(read + write) for ``bench.py``'s ``roofline.traffic``.
"""
import csv
import json
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def main():
    rep = sys.argv[1]
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    names = [short(r[kn]) for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    traffic = {}
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        cells = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.6g}"
            except ValueError:
                pass
            cells.append(v)
        print(f"| {m} ({units[i]}) | " + " | ".join(cells) + " |")
    if tj:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r, n in zip(data, names):
            b = float(r[ir].replace(",", "")) * UNIT_SCALE[units[ir]] + float(r[iw].replace(",", "")) * UNIT_SCALE[units[iw]]
            traffic.setdefault(n, []).append(b)
        with open(tj, "w") as f:
            json.dump({k: (v[0] if len(v) == 1 else v) for k, v in traffic.items()}, f, indent=1)


if __name__ == "__main__":
    main()
