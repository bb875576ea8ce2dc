#!/usr/bin/env python
"""Print the per-launch device times of an `ncu --metrics gpu__time_duration.sum --csv` log (cold-cache, serialised:
compare shares, not absolutes). usage: launch_table.py launches.csv [first_kernel_regex]"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
out = []
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("dartgpu::", "").replace("<unnamed>::", "").replace("void cub::", "cub::")[:40]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    out.append((name, v))
start = 0
if len(sys.argv) > 2:
    idx = [i for i, (n, _) in enumerate(out) if re.search(sys.argv[2], n)]
    start = idx[-1] if idx else 0
tot = sum(v for _, v in out[start:])
for n, v in out[start:]:
    print(f"{n:42s} {v:10.1f} us {100 * v / tot:5.1f}%")
print(f"{'total':42s} {tot:10.1f} us")
