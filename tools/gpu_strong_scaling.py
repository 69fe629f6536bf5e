#!/usr/bin/env python
"""GPU box (gpurun --gpus 8): gap-fill wall time of ONE draft on 1/2/4/8 GPUs through the product's in-process sharding
(FIGBIRD_GPUS=0,...: one model, gaps of the draft sharded cost-balanced, no collective) -- the second half of BASELINE's metric.
usage: tools/gpu_strong_scaling.py [workload] [max_gpus]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4s"
maxg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
os.environ["FIGBIRD_QUIET"] = "1"
case = bench.prepare_case("/tmp/fb_bench/%s_rank0" % wl, bench.WORKLOADS[wl], 102)
out = []
n = 1
while n <= maxg:
    os.environ["FIGBIRD_GPUS"] = ",".join(str(i) for i in range(n))
    ms = []
    bench.run_step_ours(case, "/tmp/fb_bench/work_strong", ms)          # warm-up: contexts, arenas
    t0 = time.perf_counter()
    steps = 2
    for _ in range(steps):
        bench.run_step_ours(case, "/tmp/fb_bench/work_strong", ms)
    dt = (time.perf_counter() - t0) / steps
    m = ms[-1]
    out.append({"gpus": n, "wall_s_per_step": dt, "device_ms_per_step_max_gpu": m["device_ms"], "t_model": m["t_model"], "t_workers": m["t_workers"], "ref_placements_p1": m["ref_placements_p1"]})
    print(json.dumps(out[-1]), flush=True)
    n *= 2
base = out[0]["wall_s_per_step"]
print(json.dumps({"workload": wl, "speedup": {str(o["gpus"]): base / o["wall_s_per_step"] for o in out}}))
