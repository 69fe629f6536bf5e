#!/usr/bin/env python
"""GPU box: fresh seeded cases in the regimes of BASELINE configs[3] / configs[4] (2x150 @ 500, gaps up to 2 kbp, negative
overlaps, reads with N, the 3500 bp library), product (CUDA engine) against the reference binaries run live.
usage: tools/gpu_live_check.py [name ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fbcase as fc  # noqa: E402

CASES = {
    "L3": ({"genome": 120000, "gaps": 12, "gapmin": 300, "gapmax": 2000, "seed": 53, "cov": 40, "sd": 50}, 150, 500),
    "L4": ({"genome": 120000, "gaps": 14, "gapmin": 10, "gapmax": 1000, "seed": 54, "cov": 40, "sd": 50, "negfrac": 0.3, "readN": 40}, 150, 500),
    "L5": ({"genome": 100000, "gaps": 10, "gapmin": 400, "gapmax": 700, "seed": 55, "cov": 50, "sd": 20}, 100, 200),
    "L6": ({"genome": 400000, "gaps": 8, "gapmin": 50, "gapmax": 3000, "seed": 56, "cov": 20, "sd": 350}, 100, 3500),
    "L8": ({"genome": 700000, "gaps": 100, "gapmin": 10, "gapmax": 2000, "seed": 61, "cov": 40, "sd": 50, "negfrac": 0.05, "readN": 200}, 150, 500),
    "L7": ({"genome": 400000, "gaps": 60, "gapmin": 10, "gapmax": 2000, "seed": 57, "cov": 40, "sd": 50, "negfrac": 0.05}, 150, 500),
}


def main():
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for name in names:
        gen, rl, ins = CASES[name]
        try:
            case = fc.make_case("/tmp/fb_live_" + name, gen, readlen=rl, insert=ins)
        except RuntimeError as e:
            print(name, "case generation failed:", str(e)[-200:], flush=True)
            bad += 1
            continue
        for mode in ("partial", "unmapped"):
            t0 = time.time(); r = fc.run_reference(case, mode, threads=os.cpu_count() or 1, worker="figbird_worker_O2"); t1 = time.time()
            o = fc.run_ours(case, mode, fc.product_exe(), threads=os.cpu_count() or 1); t2 = time.time()
            ok = {f: r[f] == o[f] for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt")}
            # the reference concatenates draw.txt in worker order (and some gaps write lines without a header): compare as multisets of lines
            ok["draw"] = sorted((r["draw.txt"] or b"").split(b"\n")) == sorted((o["draw.txt"] or b"").split(b"\n"))
            bad += sum(1 for v in ok.values() if not v)
            print(name, mode, ok, "reference %.1f s, product %.1f s" % (t1 - t0, t2 - t1), flush=True)
    print("MISMATCHES" if bad else "ALL IDENTICAL")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
