#!/usr/bin/env python3
"""Summarise `ncu --page raw --csv` exports (profiles/*_raw.csv) into profiles/ncu_summary.json: for every file the launch
with the longest duration (the dominant fb_em_kernel launch of that capture) and the counters bench.py quotes beside its
live numbers (DRAM traffic per launch, issue / shared-memory / FP64 pipe utilisation).  usage: tools/ncu_summary.py <tag=csv> ..."""
import csv
import json
import os
import sys

KEYS = {
    "duration_ms": "gpu__time_duration.sum",
    "grid": "Grid Size",
    "registers": "launch__registers_per_thread",
    "smem_dynamic_kb": "launch__shared_mem_per_block_dynamic",
    "dram_read_mb": "dram__bytes_read.sum",
    "dram_write_kb": "dram__bytes_write.sum",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smem_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_instructions": "smsp__inst_executed.sum",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
}
out = {}
for arg in sys.argv[1:]:
    tag, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    d = hdr.index(KEYS["duration_ms"])
    scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[d], 1.0)
    for r in body:
        r[d] = repr(float(r[d]) * scale)          # durations in ms whatever unit ncu chose for the column
    units[d] = "ms"
    bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for name, tgt in ((KEYS["dram_read_mb"], 1.0), (KEYS["dram_write_kb"], 1e3)):
        if name in hdr:
            i = hdr.index(name); f = bscale.get(units[i], 1.0) * tgt
            for r in body:
                r[i] = repr(float(r[i]) * f)
            units[i] = "Mbyte" if tgt == 1.0 else "Kbyte"
    best = max(body, key=lambda r: float(r[d]))
    rec = {"file": os.path.basename(path), "kernel": best[hdr.index("Kernel Name")]}
    for k, name in KEYS.items():
        if name in hdr:
            v = best[hdr.index(name)]
            try:
                rec[k] = float(v)
            except ValueError:
                rec[k] = v
            rec.setdefault("units", {})[k] = units[hdr.index(name)]
    rec["dram_bytes_per_launch"] = rec.get("dram_read_mb", 0) * 1e6 + rec.get("dram_write_kb", 0) * 1e3
    out[tag] = rec
dst = os.path.join(os.path.dirname(__file__), "..", "profiles", "ncu_summary.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
