"""Read sharding across GPUs: contiguous read ranges, pairs kept together, results concatenated in
input order, junction counts summed by key — the partition-independent merge UpdateGlobalSJMap performs
(/root/reference/src/Mapping.cpp:567-577). No collective is needed on the data path; torch.distributed is used only
by bench.py for the barrier and the max-over-ranks of the step time."""
from __future__ import annotations

from collections import OrderedDict


def shard_bounds(n_reads: int, world: int, paired: bool) -> list[int]:
    """world+1 boundaries of the contiguous read ranges; mates (2i, 2i+1) never straddle a boundary."""
    b = [n_reads * r // world for r in range(world + 1)]
    if paired:
        b = [x & ~1 for x in b]
    b[-1] = n_reads
    return b


def merge_junctions(per_rank) -> "OrderedDict[tuple[int, int], int]":
    """per_rank: iterable of iterables of (g1, g2) records in read order -> counts keyed like SpliceJunctionMap."""
    out: dict = {}
    for recs in per_rank:
        for g1, g2 in recs:
            out[(g1, g2)] = out.get((g1, g2), 0) + 1
    return OrderedDict(sorted(out.items()))


def max_over_ranks(ms: float, dist=None, device="cpu") -> float:
    """Step time of the job = the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
