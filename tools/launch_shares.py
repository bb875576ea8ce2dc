#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log by kernel name: launches, total time and share.
Index-load kernels (once per process) and the INT32 microbenchmark are listed separately: shares are of the mapping steps.
usage: launch_shares.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("dartgpu::", "").replace("<unnamed>::", "").replace("void ", "")[:44]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    tot[name] += v; cnt[name] += 1
LOAD = ("k_sa_densify", "k_relayout_occ32", "k_build_ref2", "k_ktab_level", "k_int32_peak")
load = {n: v for n, v in tot.items() if n.startswith(LOAD)}
tot = {n: v for n, v in tot.items() if not n.startswith(LOAD)}
all_us = sum(tot.values())
print(f"{'kernel':46s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{n:46s} {cnt[n]:8d} {v:12.1f} {100 * v / all_us:6.1f}%")
print(f"{'total (mapping steps)':46s} {sum(cnt[n] for n in tot):8d} {all_us:12.1f}")
for n, v in sorted(load.items(), key=lambda kv: -kv[1]):
    print(f"[once per process] {n:27s} {cnt[n]:8d} {v:12.1f}")
