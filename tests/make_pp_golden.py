#!/usr/bin/env python
"""Regenerates tests/golden/pp1.tar.gz from the REFERENCE Preprocess (oracle/_ref; needs /root/reference at build time):
a small synthetic draft + the two SAM files (tools/fbgen.cpp) and the md5 of every file the reference wrote from them in
mode 1 (maxDistance = insert) and mode 2 (1.15 x insert).  gaps_<g>.sam files are hashed over their defined fields only
(see tests/test_tools_cpu.py).  Run: python tests/make_pp_golden.py"""
import glob
import hashlib
import json
import os
import shutil
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fbcase as fc  # noqa: E402

GEN = dict(genome=36000, scaffolds=2, gaps=6, gapmin=8, gapmax=700, cov=22, sd=50, seed=81, negfrac=0.2, readN=25)
READLEN, INSERT = 150, 500


def main():
    fc.build_oracle()
    assert fc.have_reference()
    work = "/tmp/fb_pp_golden"
    shutil.rmtree(work, ignore_errors=True)
    case = fc.make_case(os.path.join(work, "full"), dict(GEN), readlen=READLEN, insert=INSERT)
    pr = fc.case_params(case)
    dst = os.path.join(work, "pp1"); os.makedirs(dst)
    for f in ("draft.fa", "result1.sam", "result2.sam"):
        shutil.copy(os.path.join(case, f), os.path.join(dst, f))
    exp = {"x": {"partial": pr["x1"], "unmapped": pr["x2"]}}
    for mode in ("partial", "unmapped"):
        d = os.path.join(case, mode); e = {}
        files = ["Temp/gapInfo.txt", "Temp/stat.txt", "Temp/stat2.txt", "myout.sam"] + sorted(
            os.path.relpath(p, d) for p in glob.glob(os.path.join(d, "Gaps", "partial_gaps_*.sam" if mode == "partial" else "gaps_*.sam")))
        for f in files:
            data = open(os.path.join(d, f), "rb").read()
            if os.path.basename(f).startswith("gaps_"):
                data = b"\n".join(b"\t".join(l.split(b"\t")[:8]) for l in data.split(b"\n"))
            e[f] = hashlib.md5(data).hexdigest()
        exp[mode] = e
    # what the reference FillGaps makes of those inputs: the end-to-end expectation of "our Preprocess, then our FillGaps"
    for mode in ("partial", "unmapped"):
        r = fc.run_reference(case, mode, threads=4)
        exp["fillgaps_" + mode] = {f: hashlib.md5(r[f]).hexdigest() for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt")}
        exp["fillgaps_" + mode]["draw.txt (per gap, sorted)"] = hashlib.md5(repr(sorted(fc.draw_by_gap(r["draw.txt"]).items())).encode()).hexdigest()
    exp["readlen"], exp["insert"] = READLEN, INSERT
    json.dump(exp, open(os.path.join(dst, "expected.json"), "w"), indent=1, sort_keys=True)
    out = os.path.join(HERE, "golden", "pp1.tar.gz")
    with tarfile.open(out, "w:gz") as t:
        t.add(dst, arcname="pp1")
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()
