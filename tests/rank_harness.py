"""Run under torchrun with the gloo backend (tests/test_ranks_cpu.py): the N > 1 protocol of bench.py without the GPUs -- rank 0 works
between host-side barriers while the other ranks wait there, every rank times the region, the maximum over ranks is reported."""
import datetime
import json
import os
import sys
import time

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figbird_b200.ranks import rank_info, reduce_counters  # noqa: E402

rank, world, local = rank_info()
dist.init_process_group("gloo")
host_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=5))      # bench.py: barriers that do not spin on a GPU
dist.barrier(group=host_group)
t0 = time.perf_counter()
steps = 3
for _ in range(steps):
    if rank == 0:
        time.sleep(0.15)                      # rank 0 fills the draft on all GPUs; the others only meet it at the barrier
    dist.barrier(group=host_group)
dt = time.perf_counter() - t0
mx, sm = reduce_counters({"dt": dt, "rank": float(rank), "work": 1000.0 * (rank + 1)}, dist, "cpu")
if rank == 0:
    print(json.dumps({"world": world, "max": mx, "sum": sm, "dt_rank0": dt, "steps": steps}))
dist.destroy_process_group()
