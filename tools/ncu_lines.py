#!/usr/bin/env python
"""Rank CUDA source lines of one kernel in an .ncu-rep by executed warp instructions and by stall samples.
usage: ncu_lines.py rep kernel_regex [launch_skip] [top]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
c, s, src = Counter(), Counter(), {}
fname, ii, ws = None, None, None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        ii, ws = r.index("Instructions Executed"), r.index("# Samples")
    elif ii and r and r[0].isdigit() and len(r) > ii and r[ii].isdigit():
        k = (fname, int(r[0]))
        c[k] += int(r[ii]); s[k] += int(r[ws]) if r[ws].isdigit() else 0; src[k] = r[1].strip()
ti, ts = sum(c.values()), sum(s.values())
print(f"total warp instructions {ti}, stall samples {ts}")
print("-- by instructions")
for k, n in c.most_common(top):
    print(f"{k[0][:20]:20s} {k[1]:4d} inst {100*n/ti:5.1f}%  smp {100*s[k]/max(ts,1):5.1f}%  {src[k][:100]}")
print("-- by stall samples")
for k, n in s.most_common(top // 2):
    print(f"{k[0][:20]:20s} {k[1]:4d} inst {100*c[k]/ti:5.1f}%  smp {100*n/max(ts,1):5.1f}%  {src[k][:100]}")
