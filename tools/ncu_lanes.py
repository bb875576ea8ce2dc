#!/usr/bin/env python
"""Per CUDA source line of one kernel in an .ncu-rep: executed warp instructions and the average number of active lanes.
usage: ncu_lanes.py rep kernel_regex [launch_skip] [top]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
inst, thr, src = Counter(), Counter(), {}
fname = None
ii = ti = None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        ii, ti = r.index("Instructions Executed"), r.index("Thread Instructions Executed")
    elif ii and r and r[0].isdigit() and len(r) > ti and r[ii].isdigit():
        k = (fname, int(r[0]))
        inst[k] += int(r[ii]); thr[k] += int(r[ti]) if r[ti].isdigit() else 0; src[k] = r[1].strip()
tot_i, tot_t = sum(inst.values()), sum(thr.values())
print(f"warp instructions {tot_i}, thread instructions {tot_t}, average active lanes {tot_t / max(tot_i, 1):.1f}")
for k, n in inst.most_common(top):
    print(f"{k[0][:18]:18s} {k[1]:4d} inst {100 * n / tot_i:5.1f}%  lanes {thr[k] / max(n, 1):5.1f}  {src[k][:110]}")
