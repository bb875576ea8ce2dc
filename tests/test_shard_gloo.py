"""The N>1 host logic on CPU: world_size 2 over gloo — shard bounds, input-order merge, junction merge, max timing."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dart_b200.shard import max_over_ranks, merge_junctions, shard_bounds


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_reads, paired, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = shard_bounds(n_reads, world, paired)
    lo, hi = b[rank], b[rank + 1]
    mine = list(range(lo, hi))                                  # stands for the per-read results of this rank
    junc = [(r // 7, r // 7 + 100) for r in mine if r % 3 == 0]  # a junction record for every third read
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, junc))
    t = max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        q.put((gathered, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    n_reads, world = 1001 * 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_reads, True, q)) for r in range(world)]
    for p in ps:
        p.start()
    gathered, t = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = [x for part, _ in gathered for x in part]
    assert merged == list(range(n_reads))                        # every read once, in input order
    assert all(len(part) % 2 == 0 for part, _ in gathered)       # mates stay together
    single = merge_junctions([[(r // 7, r // 7 + 100) for r in range(n_reads) if r % 3 == 0]])
    assert merge_junctions([j for _, j in gathered]) == single   # partition-independent junction counts
    assert t == 11.0                                             # the slowest rank defines the step


def test_shard_bounds_edges():
    assert shard_bounds(0, 4, True) == [0, 0, 0, 0, 0]
    assert shard_bounds(10, 4, True) == [0, 2, 4, 6, 10]
    assert shard_bounds(7, 2, False) == [0, 3, 7]
    for n in (2, 6, 1000, 123456):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w, True)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:])) and all(x % 2 == 0 for x in b[:-1])
