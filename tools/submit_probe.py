#!/usr/bin/env python
"""Host-side cost of dartgpu_submit / dartgpu_wait with K contexts on one host thread (not a bench number)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dart_b200 import capi
import torch
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bench.PARTS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g, idx = bench.prepare_genome()
batch = bench.as_batch(*bench.make_pairs(g, 1000000, 0))
lanes = bench.make_lanes(capi, idx, 0, dict(pair_end=1), batch, K)
lanes.upload()
for resident in (True, False):
    lanes.run(resident, 3)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); lanes.run(resident, 10); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    st = lanes.stats()
    print(f"K={K} PARTS={bench.PARTS} TURNS={os.environ.get('DARTGPU_TURNS', 'default')} resident={resident} step {dt*1e3:.2f} ms  ms_submit/step {st['ms_submit']:.3f}")
