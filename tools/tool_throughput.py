#!/usr/bin/env python
"""FASTQ -> SAM wall-clock rate of dart_b200_map on BASELINE config[1] files (1 M pairs 2x101), next to dart_ref on the same
files. usage: tool_throughput.py [pairs] [extra dart_b200_map args...]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dart_b200 import synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
extra = sys.argv[2:]
d = "/dev/shm/dart_tool" if os.path.isdir("/dev/shm") else "/tmp/dart_tool"
os.makedirs(d, exist_ok=True)
g, idx = bench.prepare_genome()
r1, r2 = os.path.join(d, f"r1_{pairs}.fq"), os.path.join(d, f"r2_{pairs}.fq")
if not os.path.exists(r2):
    with open(r1, "wb") as f1, open(r2, "wb") as f2:
        done = 0
        while done < pairs:
            k = min(1_000_000, pairs - done)
            m1, m2 = bench.make_pairs(g, k, done // 1_000_000)
            f1.write(synth.fastq_bytes(m1, 1, first_id=done)); f2.write(synth.fastq_bytes(m2, 2, first_id=done))
            done += k
tool = os.path.join(ROOT, "dart_b200", "dart_b200_map")
for rep in range(3):
    t0 = time.perf_counter()
    p = subprocess.run([tool, "-i", idx, "-f", r1, "-f2", r2, "-o", os.path.join(d, "gpu.sam"), "-j", os.path.join(d, "gpu.junc")] + extra,
                       capture_output=True, text=True)
    wall = time.perf_counter() - t0
    print("wall %.2f s |" % wall, p.stdout.split("\n")[0], flush=True)
    if p.returncode:
        print(p.stderr[-2000:])
if "-stats" in extra:
    print(p.stderr[-3000:])
print("sizes: fastq %.0f MB  sam %.0f MB" % ((os.path.getsize(r1) + os.path.getsize(r2)) / 1e6, os.path.getsize(os.path.join(d, "gpu.sam")) / 1e6))
