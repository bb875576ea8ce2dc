#!/usr/bin/env python
"""A short, fixed invocation of the hot path for ncu (never a bench number): N pairs of BASELINE config[1],
`iters` whole-path steps through the C-ABI. Usage: profile_step.py [pairs] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dart_b200 import capi  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g, idx = bench.prepare_genome()
params = dict(pair_end=1)
if bench.MIS:
    params["max_mismatch"] = int(bench.MIS)
if bench.WORKLOAD == "c5":
    params.update(multi_hit=1, max_dup=10000, all_sj=1)
M = capi.Mapper(idx, device=0, **params)
batch = bench.as_batch(*bench.make_pairs(g, pairs, 0))
for _ in range(iters):
    M.map_reads(batch, copy=False)
st = M.stats()
print({k: st[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_host", "kernel_launches")})
M.close()
