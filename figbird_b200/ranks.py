"""One process per GPU: rank discovery and the only cross-rank step of the path -- summing / maximising the
per-rank counters of a run.  Gaps are independent, so nothing else is exchanged: every rank
fills its own shard, and the host that wants whole-job figures reduces a handful of scalars."""
import os


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def reduce_counters(values, dist=None, device="cpu"):
    """values: dict name -> float of this rank.  Returns (max_over_ranks, sum_over_ranks) as dicts.
    dist: an initialised torch.distributed module (nccl on GPUs, gloo in the CPU tests) or None for one rank."""
    if dist is None:
        return dict(values), dict(values)
    import torch
    keys = sorted(values)
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64, device=device)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return dict(zip(keys, mx.tolist())), dict(zip(keys, sm.tolist()))


def shard_gaps(costs, world):
    """Longest-processing-time-first assignment of gaps to ranks by cost estimate (the same rule fb_fillgaps_main uses
    for the GPUs of one process, fb_fillgaps.cpp): returns a list of gap-index lists, one per rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for g in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(g); load[r] += costs[g] + 1
    return out
