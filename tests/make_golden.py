#!/usr/bin/env python
"""Regenerates tests/golden/*.tar.gz from the REFERENCE's own binaries (oracle/_ref; needs /root/reference at
build time).  Each fixture holds the inputs of FillGaps (as written by the reference Preprocess from fbgen's
synthetic SAM) and what the reference FillGaps + Figbird worker produced from them:

  expected/<mode>/{gapout.txt,filledContigs.fa,Ncount.txt,draw.txt}   byte-exact outputs (1 worker thread)
  expected/<mode>/model.txt        learned tables dumped by the instrumented worker (oracle/dump_patch.awk)
  expected/<mode>/counts.txt       a sample of the worker's countsGap dumps: (gap, Lg, round) -> gap rows %.17g

Run:  python tests/make_golden.py [names]   (minutes per fixture; commits nothing by itself)
"""
import os
import shutil
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fbcase as fc  # noqa: E402

CASES = {
    "g1": dict(gen={"genome": 24000, "gaplist": "12,45,95,170", "seed": 21, "cov": 30}, readlen=100, insert=200),
    "g2": dict(gen={"genome": 44000, "gaplist": "8,30,250,460,25", "seed": 22, "cov": 30, "negfrac": 0.3, "readN": 30}, readlen=100, insert=200),
    # BASELINE configs[3] parameters (2x150 bp at 500 bp insert) at fixture size
    "g3": dict(gen={"genome": 40000, "gaplist": "15,90,260", "seed": 23, "cov": 30, "sd": 50}, readlen=150, insert=500),
    # BASELINE configs[2]'s second library (3500 bp insert, maxDistance 4025): every offset of the window is admissible
    "g4": dict(gen={"genome": 50000, "gaplist": "20,110", "seed": 24, "cov": 20, "sd": 350}, readlen=100, insert=3500),
    # large gaps (N-run > unm_limit 400: one candidate, host-driven rounds with border updates, Figbird.cpp:4029-4376) at
    # BASELINE configs[3] read parameters
    "g5": dict(gen={"genome": 60000, "gaplist": "1200,2500,800", "seed": 51, "cov": 30, "sd": 50}, readlen=150, insert=500),
    # BASELINE configs[4] regime: gaps of several kbp under the 3500 bp jump library
    # 150-base reads with right-side pile-ups of more than 105 bases: update_partial_prob overruns partial_right into the
    # saved-read indices that finalize() reads (Figbird.cpp:1625-1627, 2063-2084, 5345); the host reproduces that layout
    "g7": dict(gen={"genome": 45000, "gaplist": "150,230,118", "seed": 58, "cov": 40, "sd": 50, "readN": 10}, readlen=150, insert=500),
    # scaffold ends: the reference clips its window there (initialize_start_end, Figbird.cpp:2268-2296).  Three scaffolds whose
    # first gap starts 120 / 300 / 3000 bases from the left end and whose last gap (N-run > 400: one candidate length in both
    # modes) ends read length + 10 bases before the right end; g8 under the 200 bp library (maxDistance 200 / 230: the first is
    # clipped on the left, every last gap on the right), g9 under the 3500 bp library (3500 / 4025: all of them clipped)
    "g8": dict(gen={"genome": 36000, "scaffolds": 3, "gaplist": "45,520,150,560,25,610", "gappos": "120,-110,300,-110,3000,-110", "seed": 61, "cov": 30}, readlen=100, insert=200),
    "g9": dict(gen={"genome": 60000, "scaffolds": 3, "minsep": 200, "gaplist": "60,540,130,600,35,520", "gappos": "120,-110,300,-110,3000,-110", "seed": 62, "cov": 20, "sd": 350}, readlen=100, insert=3500),
    "g6": dict(gen={"genome": 90000, "gaplist": "5000,3000,1500", "seed": 52, "cov": 20, "sd": 350}, readlen=100, insert=3500),
}


def keep_record(Lg, rnd, og):
    return (Lg % 40 == 0 or Lg == og) and rnd in (0, 2, 6)


def sample_counts(dump_path, gapinfo, out_path):
    og = {i: int(l.split()[2]) for i, l in enumerate(open(gapinfo))}
    gs = {i: int(l.split()[1]) for i, l in enumerate(open(gapinfo))}
    seen = set()
    with open(dump_path) as f, open(out_path, "w") as o:
        keep, rows_left = False, 0
        for line in f:
            if line.startswith("CALL"):
                t = line.split()
                gap, Lg, rnd = int(t[3]), int(t[5]), int(t[7])
                key = (gap, Lg, rnd)
                keep = keep_record(Lg, rnd, og[gap]) and key not in seen
                if keep:
                    seen.add(key)
                    o.write("CALL gapstart %d Lg %d round %d\n" % (gs[gap], Lg, rnd))
            elif keep:
                o.write(line)


def main():
    fc.build_oracle()
    assert fc.have_reference(), "oracle/_ref missing: /root/reference is needed to regenerate goldens"
    outdir = os.path.join(HERE, "golden")
    os.makedirs(outdir, exist_ok=True)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        case = os.path.join("/tmp", "fb_golden_" + name)
        shutil.rmtree(case, ignore_errors=True)
        fc.make_case(case, spec["gen"], readlen=spec["readlen"], insert=spec["insert"])
        for mode in ("partial", "unmapped"):
            exp = os.path.join(case, "expected", mode)
            os.makedirs(exp)
            r = fc.run_reference(case, mode, threads=1, worker="figbird_worker_O0")      # as shipped (-O0)
            r2 = fc.run_reference(case, mode, threads=1, worker="figbird_worker_O2")
            assert all(r[f] == r2[f] for f in fc.OUTPUT_FILES), "reference -O0 and -O2 disagree"
            for f in fc.OUTPUT_FILES:
                open(os.path.join(exp, f), "wb").write(r[f])
            dump = os.path.join(case, "dump_%s.txt" % mode)
            w = fc.run_reference_worker(case, mode, worker="figbird_worker_dump", extra_env={"FB_DUMP": dump, "FB_DUMP_MODEL": os.path.join(exp, "model.txt")})
            # the instrumented worker must still write the same gapout
            ref_lines = [l for l in r["gapout.txt"].split(b"\n") if l]
            assert [l for l in w["gapout"].split(b"\n") if l] == ref_lines, "dump patch changed the worker's output"
            sample_counts(dump, os.path.join(case, mode, "Temp", "gapInfo.txt"), os.path.join(exp, "counts.txt"))
            os.remove(dump)
        tar = os.path.join(outdir, name + ".tar.gz")
        with tarfile.open(tar, "w:gz") as t:
            for rel in ("draft.fa", "params.txt", "truth_gaps.txt", "partial", "unmapped", "expected"):
                t.add(os.path.join(case, rel), arcname=os.path.join(name, rel))
        print(name, "->", tar, "%.0f KB" % (os.path.getsize(tar) / 1024))


if __name__ == "__main__":
    main()
