#!/usr/bin/env python
"""Differential fuzzing of fb_preprocess_main against the reference Preprocess (oracle/_ref): the SAM of the pp1 fixture with random,
bowtie2-plausible edits per pair (flags of every pairing class, strands, positions moved next to gaps, CIGARs with soft clips /
insertions / deletions, mates on another scaffold, N-rich reads, duplicates), both modes, several maxDistance values.
usage: tools/fuzz_preprocess.py [rounds] [seed]     -> prints one line per round, exits 1 on the first difference (inputs kept).
FUZZ_MULTI=1 also repeats alignment lines (several alignments per read, as bowtie2 -k would write them)."""
import os
import random
import shutil
import subprocess
import sys
import tarfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fbcase as fc  # noqa: E402
import test_tools_cpu as tt  # noqa: E402


def rand_cigar(rng, L):
    k = rng.random()
    if k < 0.35:
        return "%dM" % L
    if k < 0.55:
        a = rng.randrange(1, L - 20); return "%dS%dM" % (a, L - a)
    if k < 0.75:
        a = rng.randrange(20, L - 1); return "%dM%dS" % (a, L - a)
    if k < 0.85:
        a = rng.randrange(1, 20); b = rng.randrange(1, 20); return "%dS%dM%dS" % (a, L - a - b, b)
    if k < 0.92:
        a = rng.randrange(10, L - 30); i = rng.randrange(1, 15); return "%dM%dI%dM" % (a, i, L - a - i)
    if k < 0.96:
        a = rng.randrange(10, L - 30); return "%dM%dD%dM" % (a, rng.randrange(1, 9), L - a)
    a = rng.randrange(5, L - 40); i = rng.randrange(1, 9); b = rng.randrange(1, 12); return "%dM%dI%dM%dS" % (a, i, L - a - i - b, b)


def mutate(rng, lines, gaps, ncontig):
    out, i = [], 0
    while i + 1 < len(lines):
        if lines[i].startswith("@") or not lines[i]:
            out.append(lines[i]); i += 1; continue
        a, b = lines[i].split("\t"), lines[i + 1].split("\t")
        i += 2
        r = rng.random()
        if r < 0.55:
            out += ["\t".join(a), "\t".join(b)]; continue
        L1, L2 = len(a[9]), len(b[9])
        s1, s2 = rng.choice((0, 16)), rng.choice((0, 16))
        kind = rng.random()
        g = rng.choice(gaps)
        near = lambda: str(max(1, g[1] + rng.randrange(-700, 700 + g[2])))      # noqa: E731
        if kind < 0.25:            # proper pair next to a gap with arbitrary CIGARs
            a[1] = str(1 | 2 | 64 | s1 | (32 if s2 else 0)); b[1] = str(1 | 2 | 128 | s2 | (32 if s1 else 0))
            a[2] = b[2] = "scaf%d" % g[0] if False else a[2]
            a[3], b[3] = near(), near(); a[5], b[5] = rand_cigar(rng, L1), rand_cigar(rng, L2)
        elif kind < 0.50:          # mate 1 aligned, mate 2 not
            a[1] = str(1 | 8 | 64 | s1); b[1] = str(1 | 4 | 128 | (32 if s1 else 0))
            a[3] = b[3] = near(); a[5] = rand_cigar(rng, L1); b[5] = "*"
        elif kind < 0.62:          # mate 1 not aligned, mate 2 aligned
            a[1] = str(1 | 4 | 64 | (32 if s2 else 0)); b[1] = str(1 | 8 | 128 | s2)
            a[3] = b[3] = near(); b[5] = rand_cigar(rng, L2); a[5] = "*"
        elif kind < 0.80:          # both aligned, not as a proper pair (too far apart, or on different scaffolds)
            a[1] = str(1 | 64 | s1 | (32 if s2 else 0)); b[1] = str(1 | 128 | s2 | (32 if s1 else 0))
            a[3], b[3] = near(), near(); a[5], b[5] = rand_cigar(rng, L1), rand_cigar(rng, L2)
            if rng.random() < 0.3 and ncontig:
                b[2] = rng.choice(ncontig)
        elif kind < 0.86:          # neither aligned
            a[1], b[1] = "77", "141"; a[2] = b[2] = "*"; a[3] = b[3] = "0"; a[5] = b[5] = "*"
        elif kind < 0.93:          # N-rich reads
            a[9] = "N" * (L1 - rng.randrange(0, 40)) + a[9][:0] + a[9][L1 - 0:]
            a[9] = (a[9] + "ACGT" * 50)[:L1]
            b[9] = ("N" * rng.randrange(0, L2) + b[9])[:L2]
        else:                      # a duplicate of the previous pair under another name
            if len(out) >= 2 and not out[-1].startswith("@"):
                pa, pb = out[-2].split("\t"), out[-1].split("\t")
                a = [a[0]] + pa[1:]; b = [b[0]] + pb[1:]
        for f in (a, b):      # bowtie2 writes MD for every aligned read (without it the reference prints uninitialised memory)
            if not (int(f[1]) & 4) and not any(t.startswith("MD:Z:") for t in f[11:]):
                f.append("MD:Z:%d" % len(f[9]))
        out += ["\t".join(a), "\t".join(b)]
        if os.environ.get("FUZZ_MULTI") and rng.random() < 0.04:      # several alignment lines per read (bowtie2 -k): the sequential path
            if int(a[1]) & 2:
                out += ["\t".join(a), "\t".join(b)]
            else:
                out.insert(len(out) - 1, "\t".join(a[:3] + [str(max(1, int(a[3]) + rng.randrange(-300, 300)))] + a[4:]))
    return out


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    work = "/tmp/fb_fuzz_pp"
    shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "pp1.tar.gz")) as t:
        t.extractall(work, filter="data")
    src = os.path.join(work, "pp1")
    draft = open(os.path.join(src, "draft.fa")).read().split("\n")
    names = [l[1:].split()[0] for l in draft if l.startswith(">")]
    # gaps of the draft: (scaffold index, start, length)
    gaps, si = [], -1
    for l in draft:
        if l.startswith(">"):
            si += 1; pos = 0; continue
        j = 0
        while j < len(l):
            if l[j] == "N":
                k = j
                while k < len(l) and l[k] == "N":
                    k += 1
                gaps.append((si, pos + j, k - j)); j = k
            else:
                j += 1
        pos += len(l)
    for rnd in range(rounds):
        rng = random.Random(seed0 * 1000 + rnd)
        case = os.path.join(work, "case%d" % rnd)
        os.makedirs(case)
        shutil.copy(os.path.join(src, "draft.fa"), os.path.join(case, "draft.fa"))
        for sam in ("result1.sam", "result2.sam"):
            lines = open(os.path.join(src, sam)).read().split("\n")
            open(os.path.join(case, sam), "w").write("\n".join(mutate(rng, lines, gaps, names)) + "\n")
        ok = True
        for mode in ("partial", "unmapped"):
            for x in ((500, 200) if mode == "partial" else (575, 230, 4025)):
                ref = os.path.join(case, "ref_%s_%d" % (mode, x)); ours = os.path.join(case, "ours_%s_%d" % (mode, x))
                try:
                    tt.run_ref_preprocess(case, mode, ref, x)
                except AssertionError:
                    print("round %d %s maxDistance %d: the reference itself failed on this input (skipped)" % (rnd, mode, x)); continue
                env = rng.choice(({"FIGBIRD_PP_BLOCK": "20000"}, {"FIGBIRD_PP_BLOCK": "2500", "FIGBIRD_HOST_THREADS": "3"}, {"FIGBIRD_PP_SEQUENTIAL": "1"}))
                tt.run_ours_preprocess(case, mode, ours, x, env=env)
                try:
                    tt.compare_dirs(ref, ours, mode)
                except AssertionError as e:
                    print("round %d %s maxDistance %d %s: DIFFERENT: %s" % (rnd, mode, x, env, str(e)[:300])); ok = False
        print("round %d %s" % (rnd, "identical" if ok else "DIFFERENT (inputs kept in %s)" % case), flush=True)
        if not ok:
            return 1
        shutil.rmtree(case, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
