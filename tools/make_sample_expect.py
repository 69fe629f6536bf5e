#!/usr/bin/env python
"""Reference results for a seeded sample of the gaps of a bench workload, computed HERE by the reference worker (oracle/_ref,
-O2 flavour -- outputs are identical to -O0, asserted by the golden fixtures) and committed as tests/golden/<workload>_sample.json:
{mode: {gapNo: gapout line}}.  bench.py compares the product's gapout.txt of its last step with it (`parity_sample` in the JSON
line) -- the at-size parity check of BASELINE configs[2] / [3], whose full reference run does not fit any budget.
The case is regenerated from its seed wherever it is needed (tools/fbgen.cpp is deterministic), so nothing but this small file
travels.  usage: python tools/make_sample_expect.py <workload> <number of gaps | all> [processes]"""
import json
import os
import random
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import fbcase as fc  # noqa: E402


def sample_gaps(workload, ngaps, k):
    return sorted(random.Random(bench.SEEDS[workload] * 7919 + 1).sample(range(ngaps), k)) if k < ngaps else list(range(ngaps))


def main():
    workload = sys.argv[1]
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    base = os.environ.get("FB_BENCH_DIR", "/tmp/fb_bench")
    case = bench.prepare_case(os.path.join(base, workload), bench.WORKLOADS[workload], bench.SEEDS[workload])
    ngaps = bench.gaps_of(case)
    k = ngaps if sys.argv[2] == "all" else min(ngaps, int(sys.argv[2]))
    chosen = sample_gaps(workload, ngaps, k)
    out = {"workload": workload, "seed": bench.SEEDS[workload], "gaps": chosen, "worker": "figbird_worker_O2"}
    for mode in ("partial", "unmapped"):
        mcase = bench.mode_case(case, mode)
        run = os.path.join(mcase, "run_sample_" + mode)
        shutil.rmtree(run, ignore_errors=True); os.makedirs(os.path.join(run, "Temp"))
        tmp = os.path.join(run, "Temp")
        for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
            shutil.copy(os.path.join(mcase, mode, "Temp", f), os.path.join(tmp, f))
        T = max(1, min(procs, len(chosen)))
        loads = [[g for i, g in enumerate(chosen) if i % T == t] for t in range(T)]
        with open(os.path.join(tmp, "gaploads.txt"), "w") as f:
            for l in loads:
                f.write("".join("%d\t" % g for g in l) + "\n")
        a = fc.fillgaps_argv(mcase, mode, tmp)
        t0 = time.time()
        ps = []
        for t in range(T):
            argv = [os.path.join(fc.REF, "figbird_worker_O2"), a[0], a[1], a[2], a[3], a[4], a[5], str(t), str(len(loads[t])), a[7], a[8], a[9], a[10], a[11], "400", a[13], a[14]]
            ps.append(subprocess.Popen(argv, cwd=run, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        rc = [p.wait() for p in ps]
        assert not any(rc), rc
        lines = {}
        for t in range(T):
            for line in open(os.path.join(tmp, "gapout%d.txt" % t)):
                lines[int(line.split("\t")[0])] = line
        assert sorted(lines) == chosen, (len(lines), len(chosen))
        out[mode] = {str(g): lines[g] for g in chosen}
        out[mode + "_seconds"] = time.time() - t0
        print(mode, "%d gaps, %.0f s on %d processes" % (len(chosen), time.time() - t0, T), flush=True)
    dst = os.path.join(ROOT, "tests", "golden", "%s_sample.json" % workload)
    json.dump(out, open(dst, "w"), sort_keys=True)
    print(dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
