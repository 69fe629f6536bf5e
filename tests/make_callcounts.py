#!/usr/bin/env python
"""Regenerates tests/golden/callcounts.json from the REFERENCE's dump-instrumented worker (oracle/_ref/figbird_worker_dump,
oracle/dump_patch.awk): for every golden fixture and mode the sequence of EM chains the reference ran, one entry
[gapStart, Lg, placeReads calls] per chain in execution order.  A chain = the calls of one initialize(): it starts at
round 0 (run() / the candidate loop, Figbird.cpp:5913-5965, 6298-6352).  tests/test_oracle_golden.py checks that the host
replay runs every one of these chains with the same number of calls (the comp_count stop rule incl. the previous_str
carry-over, Figbird.cpp:3919-3927).  Also tests/golden/placements.json: the reference's own count of pass-1 (read, offset)
placements per fixture and mode, from the counter-instrumented worker (oracle/count_patch.awk: Figbird.cpp:3169,3236,
3591,3658) -- the unit of bench.py's metric; the host replay's `ref_placements_p1` must equal it.

Run:  python tests/make_callcounts.py [names]     (a few minutes; with names only those fixtures are redone)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fbcase as fc  # noqa: E402
import golden_util as gu  # noqa: E402


def chains_from_reference_dump(path, gapinfo):
    gs = {i: int(l.split()[1]) for i, l in enumerate(open(gapinfo))}
    out = []
    with open(path) as f:
        for line in f:
            if not line.startswith("CALL"):
                continue
            t = line.split()
            gap, Lg, rnd = int(t[3]), int(t[5]), int(t[7])
            # a new chain starts at round 0, except the extra pass of a chain that stopped after its first call
            # (placeReads(i=0, finalize_flag=1) right after round 0 with finalize_flag=0, Figbird.cpp:6348-6352)
            fin = int(t[9])
            if rnd == 0 and not (out and out[-1][0] == gs[gap] and out[-1][1] == Lg and out[-1][2] == 1 and out[-1][3] == 0 and fin == 1):
                out.append([gs[gap], Lg, 1, fin])
            else:
                out[-1][2] += 1
    return [c[:3] for c in out]


def main():
    path = os.path.join(gu.GOLDEN, "callcounts.json")
    res = json.load(open(path)) if os.path.exists(path) and sys.argv[1:] else {}
    ppath = os.path.join(gu.GOLDEN, "placements.json")
    plc = json.load(open(ppath)) if os.path.exists(ppath) and sys.argv[1:] else {}
    for name in (sys.argv[1:] or gu.NAMES):
        with tempfile.TemporaryDirectory() as d:
            case = gu.extract(name, d)
            res[name] = {}
            for mode in ("partial", "unmapped"):
                dump = os.path.join(d, "dump_%s.txt" % mode)
                fc.run_reference_worker(case, mode, extra_env={"FB_DUMP": dump})
                res[name][mode] = chains_from_reference_dump(dump, os.path.join(case, mode, "Temp", "gapInfo.txt"))
                print(name, mode, len(res[name][mode]), "chains", flush=True)
                # the reference's own count of pass-1 placements (oracle/count_patch.awk); the counter build must write the same gapout
                cdir = os.path.join(d, "count_%s" % mode)
                os.makedirs(cdir)
                w = fc.run_reference_worker(case, mode, worker="figbird_worker_count", extra_env={"FB_COUNT_DIR": cdir})
                ref_lines = [l for l in gu.expected(case, mode)["gapout.txt"].split(b"\n") if l]
                assert [l for l in w["gapout"].split(b"\n") if l] == ref_lines, "count patch changed the worker's output"
                plc.setdefault(name, {})[mode] = sum(int(open(os.path.join(cdir, f)).read()) for f in os.listdir(cdir))
                print(name, mode, plc[name][mode], "pass-1 placements", flush=True)
    with open(path, "w") as f:
        json.dump(res, f, separators=(",", ":"))
    with open(ppath, "w") as f:
        json.dump(plc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
