#!/usr/bin/env python
"""Counts of pass-1 placements (the reference's own counters, oracle/count_patch.awk) on the seeded samples bench.py's reference
arm and cpu_baseline run on; written to tests/golden/bench_sample_counts.json so that a bench run need not recount.
usage: python tools/make_sample_counts.py [workload ...]   (default: c4)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    base = os.environ.get("FB_BENCH_DIR", "/tmp/fb_bench")
    dst = os.path.join(ROOT, "tests", "golden", "bench_sample_counts.json")
    known = json.load(open(dst)) if os.path.exists(dst) else {}
    cores = os.cpu_count() or 1
    for w in (sys.argv[1:] or ["c4"]):
        for ng, off in sorted(set(((bench.SAMPLE_GAPS[w][0], 2000), (bench.SAMPLE_GAPS[w][1], 2000)))):
            if ng >= bench.gen_of(w)["gaps"]:
                continue
            key = "%s_sample%d_seed%d" % (w, ng, bench.SEEDS[w] + off)
            case = bench.prepare_case(os.path.join(base, "%s_sample%d" % (w, ng)), bench.sample_spec(w, ng), bench.SEEDS[w] + off)
            known[key] = bench.count_reference_placements(case, cores)
            print(key, known[key], flush=True)
            json.dump(known, open(dst, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
