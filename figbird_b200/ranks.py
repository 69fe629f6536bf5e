"""bench.py under torchrun (one process per GPU): rank discovery and the only cross-rank step -- the maximum / sum of a handful of
per-rank scalars (the wall time of the timed steps).  Gaps are independent, so nothing is exchanged on the data path:
rank 0 drives fb_fillgaps_main over all GPUs of the box (the product's own cost-balanced sharding, fb_fillgaps.cpp), the other
ranks hold their GPU's place in the job and meet rank 0 at host-side barriers."""
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
