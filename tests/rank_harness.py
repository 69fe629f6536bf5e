"""Run under torchrun with the gloo backend (tests/test_ranks_cpu.py): every rank contributes counters, rank 0 prints the reduction."""
import json
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figbird_b200.ranks import rank_info, reduce_counters, shard_gaps  # noqa: E402

rank, world, local = rank_info()
dist.init_process_group("gloo")
vals = {"dt": 1.0 + 0.25 * rank, "dev_ms": 100.0 * (rank + 1), "placements": 1000.0 * (rank + 1)}
mx, sm = reduce_counters(vals, dist, "cpu")
costs = [float((7 * i) % 13 + 1) for i in range(40)]
mine = shard_gaps(costs, world)[rank]
mxs, sms = reduce_counters({"n": float(len(mine)), "cost": float(sum(costs[g] for g in mine))}, dist, "cpu")
if rank == 0:
    print(json.dumps({"world": world, "max": mx, "sum": sm, "shard_n_sum": sms["n"], "shard_cost_max": mxs["cost"], "shard_cost_sum": sms["cost"]}))
dist.destroy_process_group()
