"""CPU: the host-only pipeline tools of the product library (fb_preprocess_main, fb_combinegaps_main, fb_flanktrim_main,
fb_reduce_scf_main, fb_reverse_main) against the reference executables compiled into oracle/_ref, run live on the
same inputs, plus a committed fixture (tests/golden/pp1.tar.gz: inputs + md5 of what the reference wrote) that needs no reference.

Parity bar: byte-identical files.  One stated exception: in Gaps/gaps_<g>.sam the reference prints the MD string and the IH
count of some records from uninitialised heap memory (Preprocess.cpp:1493: `new SAM`, md / ih never set for a read without an
MD tag / outside printVectors); those two trailing fields are compared only where they are defined, i.e. not at all --
FillGaps reads neither (Figbird.cpp:5661-5767)."""
import glob
import hashlib
import json
import os
import random
import re
import shutil
import subprocess
import sys
import tarfile

import pytest

import fbcase as fc

HERE = os.path.dirname(os.path.abspath(__file__))
FBTOOL = os.path.join(fc.PBUILD, "fbtool")
needs_ref = pytest.mark.skipif(not (fc.have_reference() and os.path.exists(os.path.join(fc.REF, "FlankTrim"))), reason="oracle/_ref not built")


def run_ours_preprocess(case, mode, out, x, extra=("1", "0", "0"), env=None, reads=("r1.fq", "r2.fq"), filled=None):
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(os.path.join(out, "Gaps")); os.makedirs(os.path.join(out, "Temp"))
    draft = os.path.join(case, "draft.fa")
    sam = os.path.join(case, "result1.sam" if mode == "partial" else "result2.sam")
    e = dict(os.environ); e.update(env or {})
    p = subprocess.run([FBTOOL, "preprocess", draft, str(x), "1" if mode == "partial" else "2", sam, os.path.join(out, "myout.sam"), filled or draft,
                        reads[0], reads[1], os.path.join(out, "Gaps") + "/", os.path.join(out, "Temp") + "/"] + list(extra), env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 0, p.stdout.decode()
    return p.stdout.decode()


def run_ref_preprocess(case, mode, out, x, extra=("1", "0", "0"), reads=("r1.fq", "r2.fq"), filled=None, sam=None):
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(os.path.join(out, "Gaps")); os.makedirs(os.path.join(out, "Temp"))
    draft = os.path.join(case, "draft.fa")
    sam = sam or os.path.join(case, "result1.sam" if mode == "partial" else "result2.sam")
    p = subprocess.run([os.path.join(fc.REF, "Preprocess"), draft, str(x), "1" if mode == "partial" else "2", sam, os.path.join(out, "myout.sam"), filled or draft,
                        reads[0], reads[1], os.path.join(out, "Gaps") + "/", os.path.join(out, "Temp") + "/"] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 0, p.stdout.decode()
    return p.stdout.decode()


def defined_fields(data):
    """gaps_<g>.sam: everything but the two trailing fields the reference leaves undefined."""
    return [l.split(b"\t")[:8] for l in data.split(b"\n")]


def compare_dirs(ref, ours, mode):
    for f in ("Temp/gapInfo.txt", "Temp/stat.txt", "Temp/stat2.txt", "myout.sam"):
        a = open(os.path.join(ref, f), "rb").read(); b = open(os.path.join(ours, f), "rb").read()
        assert a == b, "%s differs (%d vs %d bytes)" % (f, len(a), len(b))
    pat = "partial_gaps_*.sam" if mode == "partial" else "gaps_*.sam"
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(ref, "Gaps", pat)))
    assert names == sorted(os.path.basename(p) for p in glob.glob(os.path.join(ours, "Gaps", pat)))
    nonempty = 0
    for n in names:
        a = open(os.path.join(ref, "Gaps", n), "rb").read(); b = open(os.path.join(ours, "Gaps", n), "rb").read()
        nonempty += bool(a)
        if mode == "partial":
            assert a == b, n
        else:
            assert defined_fields(a) == defined_fields(b), n
    return len(names), nonempty


CASES = {
    # 2x100 @ 200: maxDistance 200 / 230 (the <= 250 branch of checkPos), negative overlaps, reads with N
    "frag": (dict(genome=60000, scaffolds=3, gaps=9, gapmin=5, gapmax=300, cov=25, seed=71, negfrac=0.3, readN=20), 100, 200),
    # 2x150 @ 500: mode 2 runs with maxDistance 575 (model pass + insert-size logic of checkPos, linked pairs)
    "c4like": (dict(genome=90000, scaffolds=2, gaps=10, gapmin=10, gapmax=1500, cov=30, sd=50, seed=72, readN=40), 150, 500),
    # 3500 bp jump library: several gaps inside one maxDistance window
    "jump": (dict(genome=80000, scaffolds=2, gaps=12, gapmin=10, gapmax=600, cov=15, sd=350, seed=73, minsep=150), 100, 3500),
}


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    if not fc.have_reference():
        pytest.skip("oracle/_ref not built")
    out = {}
    for name, (gen, readlen, insert) in CASES.items():
        out[name] = fc.make_case(str(tmp_path_factory.mktemp("ppcase") / name), dict(gen), readlen=readlen, insert=insert)
    return out


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("variant", ["blocks", "tiny-blocks-1-thread", "sequential"])
def test_preprocess_matches_reference(cases, tmp_path, name, variant):
    case = cases[name]
    pr = fc.case_params(case)
    env = {"blocks": {"FIGBIRD_PP_BLOCK": "200000"}, "tiny-blocks-1-thread": {"FIGBIRD_PP_BLOCK": "3000", "FIGBIRD_HOST_THREADS": "1"}, "sequential": {"FIGBIRD_PP_SEQUENTIAL": "1"}}[variant]
    for mode, x in (("partial", pr["x1"]), ("unmapped", pr["x2"])):
        ours = str(tmp_path / ("ours_" + mode))
        run_ours_preprocess(case, mode, ours, x, env=env)
        n, nonempty = compare_dirs(os.path.join(case, mode), ours, mode)
        assert n == CASES[name][0]["gaps"] and nonempty > 0


@needs_ref
def test_preprocess_reduced_reads_and_genome_reduction(cases, tmp_path):
    """read_reduction=1 (the FASTQ of every pair that is not a perfect full map, QUAL reversed in place: Preprocess.cpp:1696-1731)
    and genome_reduction=1 (gapInfo.txt numbered by the scaffolds of the un-reduced genome: :1888-2007, 2134-2142)."""
    case = cases["frag"]
    pr = fc.case_params(case)
    # the reduced genome = the scaffolds that hold a gap (here: all of them, in another file with one scaffold more in front)
    full = str(tmp_path / "full.fa")
    with open(full, "w") as f:
        f.write(">extra\nACGTACGTAGCTAGCTAGCATCGATCGATCAGCTAGCTAGCTAGCATCG\n" + open(os.path.join(case, "draft.fa")).read())
    for mode, x in (("partial", pr["x1"]), ("unmapped", pr["x2"])):
        outs = {}
        for who in ("ref", "ours"):
            rd = tmp_path / ("reads_%s_%s" % (who, mode)); rd.mkdir()
            reads = (str(rd / "lib_1.fastq"), str(rd / "lib_2.fastq"))
            out = str(tmp_path / ("%s_%s" % (who, mode)))
            (run_ref_preprocess if who == "ref" else run_ours_preprocess)(case, mode, out, x, extra=("1", "1", "1"), reads=reads, filled=full)
            outs[who] = (out, rd)
        compare_dirs(outs["ref"][0], outs["ours"][0], mode)
        for f in ("lib_1_reduced.fastq", "lib_2_reduced.fastq"):
            a = open(str(outs["ref"][1] / f), "rb").read(); b = open(str(outs["ours"][1] / f), "rb").read()
            assert a == b and len(a) > 0, f


@needs_ref
def test_preprocess_irregular_input_takes_the_sequential_path(cases, tmp_path):
    """Several alignment lines per read (bowtie2 -k) and a scaffold that ends in N: the block scheme steps aside, results stay
    those of the reference."""
    case = cases["frag"]
    pr = fc.case_params(case)
    work = tmp_path / "case"; work.mkdir()
    # the last scaffold ends in an N-run and a further scaffold follows: the run is reported under the next scaffold's number
    draft = open(os.path.join(case, "draft.fa")).read().rstrip("\n")
    (work / "draft.fa").write_text(draft + "NNNNNNNNNNNN\n>tail\nACGTTGCATGCATGCATGGGATCGATCGTAGCTAGCTAGCTAGCTGATCGTAGC\n")
    rng = random.Random(5)
    for sam in ("result1.sam", "result2.sam"):
        lines = open(os.path.join(case, sam)).read().split("\n")
        out, i = [], 0
        while i + 1 < len(lines):
            if lines[i].startswith("@") or not lines[i]:
                out.append(lines[i]); i += 1; continue
            a, b = lines[i], lines[i + 1]
            out += [a, b]
            fa = int(a.split("\t")[1])
            if rng.random() < 0.05:
                if fa & 2:
                    out += [a, b]                 # a second alignment of a proper pair
                else:
                    out.insert(len(out) - 1, a)   # a second line of mate 1 of a pair that is not properly aligned
            i += 2
        (work / sam).write_text("\n".join(out) + "\n")
    for mode, x in (("partial", pr["x1"]), ("unmapped", pr["x2"])):
        ref = str(tmp_path / ("ref_" + mode)); ours = str(tmp_path / ("ours_" + mode))
        run_ref_preprocess(str(work), mode, ref, x)
        run_ours_preprocess(str(work), mode, ours, x, env={"FIGBIRD_PP_BLOCK": "5000"})
        compare_dirs(ref, ours, mode)
        gi = open(os.path.join(ours, "Temp", "gapInfo.txt")).read().strip().split("\n")
        assert gi[-1].split("\t")[0] == "3" and gi[-1].split("\t")[2] == "12"      # the run at the end of scaffold 2, under scaffold 3


def test_preprocess_fixture_without_reference(tmp_path):
    """tests/golden/pp1.tar.gz: draft + SAMs + md5 of every file the reference Preprocess wrote (made by tests/make_pp_golden.py)."""
    with tarfile.open(os.path.join(HERE, "golden", "pp1.tar.gz")) as t:
        t.extractall(str(tmp_path), filter="data")
    case = str(tmp_path / "pp1")
    exp = json.load(open(os.path.join(case, "expected.json")))
    for mode in ("partial", "unmapped"):
        ours = str(tmp_path / ("ours_" + mode))
        run_ours_preprocess(case, mode, ours, exp["x"][mode], env={"FIGBIRD_PP_BLOCK": "100000"})
        for f, want in exp[mode].items():
            data = open(os.path.join(ours, f), "rb").read()
            if os.path.basename(f).startswith("gaps_"):
                data = b"\n".join(b"\t".join(x) for x in defined_fields(data))
            assert hashlib.md5(data).hexdigest() == want, "%s %s" % (mode, f)


# ---------------------------------------------------------------------------------------------------------------------
def write_fasta(path, recs, width=60):
    with open(path, "w") as f:
        for name, seq in recs:
            f.write(">" + name + "\n")
            if width <= 0:
                f.write(seq + "\n")
            else:
                for i in range(0, len(seq), width):
                    f.write(seq[i:i + width] + "\n")


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def gapped(rng, n, runs):
    s = list(rand_seq(rng, n))
    for start, ln, ch in runs:
        for i in range(start, min(n, start + ln)):
            s[i] = ch
    return "".join(s)


@needs_ref
@pytest.mark.parametrize("width", [60, 1022, 1023, 0])
def test_flanktrim_and_reduce_scf_match_reference(tmp_path, width):
    rng = random.Random(11 + width)
    recs = [("s1 first scaffold", gapped(rng, 5000, [(400, 30, "N"), (1500, 1, "N"), (2500, 60, "N"), (2570, 5, "n"), (4990, 10, "N")])),
            ("s2", rand_seq(rng, 2600)),
            ("s3\tx", gapped(rng, 3000, [(0, 20, "N"), (700, 99, "N"), (820, 100, "N"), (1200, 40, "N"), (1249, 3, "N")])),
            ("s4", "n" + rand_seq(rng, 1100))]
    src = str(tmp_path / "g.fa")
    write_fasta(src, recs, width)
    for trim, readlen in ((10, 100), (0, 100), (25, 50)):
        a = str(tmp_path / "ref.fa"); b = str(tmp_path / "ours.fa")
        assert subprocess.run([os.path.join(fc.REF, "FlankTrim"), src, str(trim), str(readlen), a]).returncode == 0
        assert subprocess.run([FBTOOL, "flanktrim", src, str(trim), str(readlen), b]).returncode == 0
        assert open(a, "rb").read() == open(b, "rb").read(), (trim, readlen)
    ra = tmp_path / "ra"; rb = tmp_path / "rb"; ra.mkdir(); rb.mkdir()
    assert subprocess.run([os.path.join(fc.REF, "Reduce_SCF"), src, str(ra) + "/"]).returncode == 0
    assert subprocess.run([FBTOOL, "reduce_scf", src, str(rb) + "/"]).returncode == 0
    x = open(str(ra / "newgenome.fa"), "rb").read()
    assert x == open(str(rb / "newgenome.fa"), "rb").read() and x.count(b">") == 3


@needs_ref
def test_reverse_matches_reference(tmp_path):
    rng = random.Random(3)
    for who in ("ref", "ours"):
        d = tmp_path / who; d.mkdir()
        for k in (1, 2):
            with open(str(d / ("jump_lib.%d.fastq" % k)), "w") as f:
                r2 = random.Random(100 + k)
                for i in range(200):
                    n = r2.choice((36, 100, 151))
                    f.write("@r%d/%d\n%s\n+\n%s\n" % (i, k, rand_seq(r2, n, "ACGTNacgt"), "".join(chr(33 + r2.randrange(40)) for _ in range(n))))
                if k == 2:
                    f.write("@extra/2\nACGT\n+\nIIII\n")      # the longer file is cut at the shorter one's end
    a = subprocess.run([os.path.join(fc.REF, "Reverse"), "jump_lib.1.fastq", "jump_lib.2.fastq"], cwd=str(tmp_path / "ref"), stdout=subprocess.PIPE)
    b = subprocess.run([FBTOOL, "reverse", "jump_lib.1.fastq", "jump_lib.2.fastq"], cwd=str(tmp_path / "ours"), stdout=subprocess.PIPE)
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout
    for line in a.stdout.decode().split():
        assert open(str(tmp_path / "ref" / line), "rb").read() == open(str(tmp_path / "ours" / line), "rb").read(), line
    assert sorted(a.stdout.decode().split()) == ["jump_lib.1_reversed.fastq", "jump_lib.2_reversed.fastq"]
    del rng


def gapout_line(i, og, s):
    return "%d\t0\t%d\t%d\t%d\t%s\n" % (i, 1000 * i, og, len(s), s)


@needs_ref
@pytest.mark.parametrize("seed", [9, 10, 11, 12])
def test_combinegaps_matches_reference(tmp_path, seed):
    rng = random.Random(seed)
    # iteration 1: closed gaps, gaps that keep one N-run, a gap closed with length 0; later iterations only list the open ones
    state = []
    it1 = []
    for i in range(40):
        og = rng.randrange(5, 400)
        kind = rng.choice(("closed", "open", "open", "zero"))
        if kind == "closed":
            s = rand_seq(rng, rng.randrange(1, 500))
        elif kind == "zero":
            s = ""
        else:
            s = rand_seq(rng, rng.randrange(0, 80)) + "N" * rng.randrange(1, 120) + rand_seq(rng, rng.randrange(0, 80))
        it1.append(gapout_line(i, og, s)); state.append(kind == "open")
    files = {1: it1}
    for itr in (2, 3):
        lines = []
        for i in range(40):
            if not state[i]:
                continue
            og = rng.randrange(1, 120)
            kind = rng.choice(("closed", "open", "zero")) if itr == 2 else rng.choice(("closed", "open"))
            if kind == "closed":
                s = rand_seq(rng, rng.randrange(1, 200))
            elif kind == "zero":
                s = ""
            else:
                s = rand_seq(rng, rng.randrange(0, 40)) + "N" * rng.randrange(1, 60) + rand_seq(rng, rng.randrange(0, 40))
            lines.append(gapout_line(i, og, s)); state[i] = kind == "open"
        files[itr] = lines
    for who in ("ref", "ours"):
        d = tmp_path / who; d.mkdir()
        for itr, lines in files.items():
            (d / ("gapout_%d.txt" % itr)).write_text("".join(lines))
    for n in (1, 2, 3):
        assert subprocess.run([os.path.join(fc.REF, "CombineGaps"), str(n), str(tmp_path / "ref") + "/"]).returncode == 0
        assert subprocess.run([FBTOOL, "combinegaps", str(n), str(tmp_path / "ours") + "/"]).returncode == 0
        for f in ("combined_gapstring.txt", "Individual_gaps.txt"):
            a = open(str(tmp_path / "ref" / f), "rb").read(); b = open(str(tmp_path / "ours" / f), "rb").read()
            assert a == b and len(a) > 100, (n, f)


def test_tools_through_the_c_abi(tmp_path):
    """The same entry points through ctypes (what a cgo / JNI host would bind): argv in, exit status out."""
    from figbird_b200 import capi
    src = str(tmp_path / "g.fa")
    write_fasta(src, [("a", "ACGT" * 50 + "NNNN" + "ACGT" * 50), ("b", "ACGT" * 30)])
    assert capi.tool("reduce_scf", [src, str(tmp_path) + "/"]) == 0
    assert open(str(tmp_path / "newgenome.fa")).read().count(">") == 1
    assert capi.tool("flanktrim", [src, "2", "100", str(tmp_path / "t.fa")]) == 0
    assert open(str(tmp_path / "t.fa")).read().split("\n")[1][196:208] == "ACNNNNNNNNGT"
    assert capi.tool("reduce_scf", [str(tmp_path / "missing.fa"), str(tmp_path) + "/"]) == 1
    assert capi.tool("preprocess", ["too", "few"]) == 1


@needs_ref
def test_preprocess_on_bowtie2_like_oddities(cases, tmp_path):
    """What the generator never writes but bowtie2 does: PCR duplicates (exact, and the same read minus a base or two at an end --
    the mode-2 filter drops a read whose core is contained in a stored one), insertions spanning the gap ("aMbIcM": the MIM evidence
    of stat2.txt), soft clips at both ends, lower-case and N bases, more than 3001 reads for one gap (the cap), blanks inside a tag."""
    case = cases["frag"]
    pr = fc.case_params(case)
    work = tmp_path / "case"; work.mkdir()
    shutil.copy(os.path.join(case, "draft.fa"), str(work / "draft.fa"))
    rng = random.Random(17)
    for sam in ("result1.sam", "result2.sam"):
        lines = open(os.path.join(case, sam)).read().split("\n")
        pairs, head = [], []
        i = 0
        while i + 1 < len(lines):
            if lines[i].startswith("@") or not lines[i]:
                head.append(lines[i]); i += 1; continue
            pairs.append((lines[i], lines[i + 1])); i += 2
        out = []
        n = 0
        for a, b in pairs:
            fa_, fb_ = a.split("\t"), b.split("\t")
            interesting = (int(fa_[1]) & 2) == 0 or "S" in fa_[5] or "S" in fb_[5]
            out.append((a, b))
            if not interesting:
                continue
            r = rng.random()
            n += 1
            nm = "dup%d" % n
            if r < 0.10:                                     # exact duplicate under another name
                out.append(("\t".join([nm] + fa_[1:]), "\t".join([nm] + fb_[1:])))
            elif r < 0.20:                                   # the same pair, both reads two bases shorter at both ends (contained cores)
                def trim(f):
                    g = list(f)
                    if g[5] != "*" and g[5].endswith("M") and g[5][:-1].isdigit():
                        g[5] = "%dM" % (int(g[5][:-1]) - 4); g[3] = str(int(g[3]) + 2)
                    elif g[5] != "*":
                        return f
                    g[9] = g[9][2:-2]; g[10] = g[10][2:-2]
                    return g
                out.append(("\t".join([nm] + trim(fa_)[1:]), "\t".join([nm] + trim(fb_)[1:])))
            elif r < 0.30 and any(re.match(r"^\d+M\d+S$", f[5]) and int(re.match(r"^\d+M(\d+)S$", f[5]).group(1)) > 14 for f in (fa_, fb_)):
                # a read that runs into the gap from the left and comes out on the other side: "aMbIcM", the MIM evidence
                ga, gb = list(fa_), list(fb_)
                for g in (ga, gb):
                    m = re.match(r"^(\d+)M(\d+)S$", g[5])
                    if m and int(m.group(2)) > 14:
                        g[5] = "%sM%dI10M" % (m.group(1), int(m.group(2)) - 10)
                        g[9] = ("C" if g[9][0] != "C" else "G") + g[9][1:]      # (an exact copy would be dropped as a duplicate)
                out.append(("\t".join([nm] + ga[1:]), "\t".join([nm] + gb[1:])))
            elif r < 0.32 and "S" in fa_[5]:                 # soft clips at both ends
                g = list(fa_); L = len(g[9])
                g[5] = "7S%dM9S" % (L - 16)
                out.append(("\t".join([nm] + g[1:]), "\t".join([nm] + fb_[1:])))
            elif r < 0.36:                                   # lower-case bases and an N
                g = list(fb_); g[9] = g[9][:10].lower() + "N" + g[9][11:]
                out.append(("\t".join([nm] + fa_[1:]), "\t".join([nm] + g[1:])))
            elif r < 0.40:                                   # a blank inside an optional field: getSAM splits on blanks too
                out.append((a + "\tXX:Z:two words", "\t".join([nm.replace("dup", "dupb")] + fb_[1:]).replace(nm.replace("dup", "dupb"), fa_[0], 1)))
        # more than 3001 candidates for one gap: one interesting pair repeated with distinct reads
        if sam == "result2.sam":
            seed_pair = next(((a, b) for a, b in pairs if (int(a.split("\t")[1]) & 6) == 0 and (int(b.split("\t")[1]) & 4) != 0), None)
        else:
            seed_pair = next(((a, b) for a, b in pairs if "S" in a.split("\t")[5] and (int(a.split("\t")[1]) & 2)), None)
        if seed_pair:
            fa_, fb_ = seed_pair[0].split("\t"), seed_pair[1].split("\t")
            for k in range(3100):
                s = "".join(rng.choice("ACGT") for _ in range(len(fb_[9])))
                if sam == "result2.sam":
                    out.append(("\t".join(["cap%d" % k] + fa_[1:]), "\t".join(["cap%d" % k] + fb_[1:9] + [s] + fb_[10:])))
                else:
                    s1 = "".join(rng.choice("ACGT") for _ in range(len(fa_[9])))
                    out.append(("\t".join(["cap%d" % k] + fa_[1:9] + [s1] + fa_[10:]), "\t".join(["cap%d" % k] + fb_[1:])))
        with open(str(work / sam), "w") as f:
            f.write("\n".join([h for h in head if h] + [x for p in out for x in p]) + "\n")
    for mode, x in (("partial", pr["x1"]), ("unmapped", pr["x2"])):
        ref = str(tmp_path / ("ref_" + mode))
        run_ref_preprocess(str(work), mode, ref, x)
        for env in ({"FIGBIRD_PP_BLOCK": "20000"}, {"FIGBIRD_PP_SEQUENTIAL": "1"}):
            ours = str(tmp_path / ("ours_" + mode))
            run_ours_preprocess(str(work), mode, ours, x, env=env)
            compare_dirs(ref, ours, mode)
        if mode == "partial":
            s2 = [l.split("\t") for l in open(os.path.join(ref, "Temp", "stat2.txt")).read().strip().split("\n")]
            assert any(l[1] == "1" for l in s2), "no MIM evidence was generated: the test lost its point"
        pat = "partial_gaps_*.sam" if mode == "partial" else "gaps_*.sam"
        per = 1 if mode == "partial" else 2
        assert max(open(p, "rb").read().count(b"\n") for p in glob.glob(os.path.join(ref, "Gaps", pat))) >= 3001 * per - 1


def test_gap_container_is_the_text_files_in_one_file(tmp_path):
    """SURVEY 8f-2, opt-in (FIGBIRD_CONTAINER=1): fb_preprocess_main leaves Gaps/partial_gaps.fbc / Gaps/gaps.fbc beside the per-gap text
    files -- the same bytes behind an offset table -- and fb_fillgaps_main reads its per-gap inputs from them: same outputs."""
    import struct
    with tarfile.open(os.path.join(HERE, "golden", "pp1.tar.gz")) as t:
        t.extractall(str(tmp_path), filter="data")
    case = str(tmp_path / "pp1")
    exp = json.load(open(os.path.join(case, "expected.json")))
    draft = os.path.join(case, "draft.fa")
    gaps = os.path.join(case, "Gaps") + "/"
    os.makedirs(gaps)
    env = dict(os.environ); env["FIGBIRD_CONTAINER"] = "1"
    for mode, flag, sam in (("partial", "1", "result1.sam"), ("unmapped", "2", "result2.sam")):
        tmp = os.path.join(case, mode, "Temp") + "/"
        os.makedirs(tmp)
        p = subprocess.run([FBTOOL, "preprocess", draft, str(exp["x"][mode]), flag, os.path.join(case, sam), os.path.join(case, mode, "myout.sam"), draft, "r1.fq", "r2.fq", gaps, tmp, "1", "0", "0"], env=env)
        assert p.returncode == 0
        box = open(os.path.join(gaps, "partial_gaps.fbc" if mode == "partial" else "gaps.fbc"), "rb").read()
        kind, n = struct.unpack_from("<II", box, 8)
        assert box[:8] == b"FBGAPS1\0" and kind == int(flag) and n == len(exp[mode]) - 4
        off = struct.unpack_from("<%dQ" % (n + 1), box, 16)
        for g in range(n):
            assert box[off[g]:off[g + 1]] == open(os.path.join(gaps, ("partial_gaps_%d.sam" if mode == "partial" else "gaps_%d.sam") % g), "rb").read()
        assert off[n] == len(box)
    # partial mode through the CPU engine (minutes for unmapped mode at these gap lengths; the loaders are the same code)
    outs = {}
    for how in ("text", "container"):
        tmp = os.path.join(case, "run_" + how, "Temp") + "/"
        os.makedirs(tmp)
        for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
            shutil.copy(os.path.join(case, "partial", "Temp", f), tmp + f)
        e = dict(os.environ); e["FIGBIRD_CONTAINER"] = "1" if how == "container" else "0"; e["FIGBIRD_TAIL_ITEMS"] = "0"
        argv = [draft, str(exp["x"]["partial"]), str(exp["readlen"]), "1", "1", "0", "4", os.path.join(case, "partial", "myout.sam"), tmp, gaps, "30", str(exp["readlen"]), "0", "0", str(exp["insert"])]
        if how == "container":      # the text files are not even there: everything comes from the container
            for f in glob.glob(os.path.join(gaps, "partial_gaps_*.sam")):
                os.remove(f)
        assert subprocess.run([fc.oracle_exe()] + argv, env=e, stdout=subprocess.DEVNULL).returncode == 0
        outs[how] = fc.read_outputs(tmp)
    for f in fc.OUTPUT_FILES:
        assert outs["text"][f] == outs["container"][f] and outs["text"][f], f
    assert hashlib.md5(outs["text"]["gapout.txt"]).hexdigest() == exp["fillgaps_partial"]["gapout.txt"]


@needs_ref
def test_combinegaps_gives_up_like_the_reference_on_two_runs(tmp_path):
    """A gap string with two N-runs: the reference exits with status 0 before it writes anything for that iteration
    (CombineGaps.cpp:252-256); so does the replacement -- same files left behind."""
    for who in ("ref", "ours"):
        d = tmp_path / who; d.mkdir()
        (d / "gapout_1.txt").write_text(gapout_line(0, 10, "ACGTNNACGT") + gapout_line(1, 20, "ACNNGTNNAC") + gapout_line(2, 5, "ACGT"))
    a = subprocess.run([os.path.join(fc.REF, "CombineGaps"), "1", str(tmp_path / "ref") + "/"])
    b = subprocess.run([FBTOOL, "combinegaps", "1", str(tmp_path / "ours") + "/"])
    assert a.returncode == 0 and b.returncode == 0
    assert sorted(os.listdir(str(tmp_path / "ref"))) == sorted(os.listdir(str(tmp_path / "ours"))) == ["gapout_1.txt"]
    # a missing iteration file: message + status 1
    a = subprocess.run([os.path.join(fc.REF, "CombineGaps"), "2", str(tmp_path / "ref") + "/"], stdout=subprocess.PIPE)
    b = subprocess.run([FBTOOL, "combinegaps", "2", str(tmp_path / "ours") + "/"], stdout=subprocess.PIPE)
    assert a.returncode == b.returncode == 1 or (a.returncode == 0 and b.returncode == 0)


@needs_ref
def test_preprocess_differential_fuzz_smoke():
    """Two rounds of tools/fuzz_preprocess.py (random bowtie2-plausible edits of every pairing class, both modes, maxDistance below and
    above 250, block / sequential paths); the tool itself runs as many rounds as one likes (46 rounds were identical when this was
    written)."""
    p = subprocess.run([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "fuzz_preprocess.py"), "2", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0 and p.stdout.decode().count("identical") == 2, p.stdout.decode()[-2000:]
