"""Markdown table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list: one row per distinct (kernel, grid) with the device time, the measured DRAM traffic, the achieved DRAM
GB/s and its fraction of the measured HBM copy peak (MEASURED_PEAKS.json).  Torch's own elementwise kernels (input
generation of the profiling scripts) are skipped.

    python tools/kernel_rooflines.py profiles/<launch list>.csv [...] > profiles/<name>.md
"""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6569.3


def load(path):
    txt = open(path).read()
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    hdr = rows[0]
    by = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        e = by.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    return list(by.values())


def main():
    for path in sys.argv[1:]:
        print(f"### {os.path.basename(path)}\n")
        print("| kernel | grid | block | launches | us (median) | DRAM read MB | DRAM write MB | DRAM GB/s | of %.0f GB/s |" % PEAK)
        print("|---|---|---|---|---|---|---|---|---|")
        groups = collections.OrderedDict()
        for d in load(path):
            nm = d["name"].split("(")[0].replace("ssdsb::<unnamed>::", "").replace("void ", "")
            if nm.startswith("at::"):
                continue
            groups.setdefault((nm, d["grid"], d["block"]), []).append(d)
        for (nm, grid, block), ds in groups.items():
            ds = sorted(ds, key=lambda d: d["gpu__time_duration.sum"])
            m = ds[len(ds) // 2]
            t = m["gpu__time_duration.sum"] / 1e3
            rd, wr = m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6
            gbs = (rd + wr) / t * 1e3 if t else 0            # MB / us = TB/s -> x1e3 = GB/s
            print(f"| `{nm}` | {grid} | {block} | {len(ds)} | {t:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {gbs / PEAK:.2f} |")
        print()


if __name__ == "__main__":
    main()
