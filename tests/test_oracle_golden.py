"""CPU: the oracle (our host program + CPU restatement engine) against golden vectors produced by the
reference's own binaries (tests/make_golden.py).  This is what pins the oracle.

  * gapout.txt / filledContigs.fa / Ncount.txt / draw.txt byte-identical, both modes;
  * learned model tables identical (%.17g text);
  * countsGap gap rows of sampled (gap, Lg, round) calls within 1e-9 relative of the instrumented reference.
"""
import os

import numpy as np
import pytest

import fbcase as fc
import golden_util as gu


@pytest.fixture(scope="module", params=gu.NAMES)
def case(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_" + request.param)
    return gu.extract(request.param, str(d))


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_oracle_matches_reference_outputs(case, mode):
    dump = os.path.join(case, "oracle_counts_%s.txt" % mode)
    model = os.path.join(case, "oracle_model_%s.txt" % mode)
    metrics = os.path.join(case, "oracle_metrics_%s.json" % mode)
    o = fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FB_ORACLE_DUMP": dump, "FIGBIRD_DUMP_MODEL": model, "FIGBIRD_METRICS": metrics}, name="oracle")
    exp = gu.expected(case, mode)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert o[f] == exp[f], "%s differs from the reference (%s mode)" % (f, mode)
    # pass-1 placements the reference scan consumed: the host replay's count against the reference's own counters
    # (oracle/count_patch.awk at Figbird.cpp:3169,3236,3591,3658; tests/golden/placements.json)
    import json
    want = json.load(open(os.path.join(gu.GOLDEN, "placements.json")))[os.path.basename(case)][mode]
    assert json.load(open(metrics))["ref_placements_p1"] == want
    # model tables
    assert gu.model_lines(model) == gu.model_lines(os.path.join(case, "expected", mode, "model.txt"))
    # per-position base weights
    ref = gu.parse_counts(os.path.join(case, "expected", mode, "counts.txt"))
    mine = gu.parse_counts(dump)
    common = [k for k in ref if k in mine]
    assert len(common) >= max(3, len(ref) // 2), "too few comparable countsGap records (%d of %d)" % (len(common), len(ref))
    worst = 0.0
    for k in common:
        a, b = ref[k], mine[k]
        assert a.shape == b.shape
        nz = a != 0
        assert np.array_equal(nz, b != 0)
        if nz.any():
            worst = max(worst, float(np.max(np.abs(a[nz] - b[nz]) / np.abs(a[nz]))))
    assert worst <= 1e-9, "countsGap relative error %g" % worst


def _chains(dump):
    """{(gapstart, Lg): [placeReads calls of the 1st, 2nd, ... EM chain run at that length]} from an FB_ORACLE_DUMP file."""
    out = {}
    with open(dump) as f:
        for line in f:
            if line.startswith("CALL"):
                t = line.split()
                key, rnd = (int(t[2]), int(t[4])), int(t[6])
                if rnd == 0:
                    out.setdefault(key, []).append(1)
                else:
                    out[key][-1] += 1
    return out


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_em_chain_lengths_match_reference(case, mode):
    """Every EM chain the reference ran (one initialize() + its placeReads calls, from the dump-instrumented worker:
    tests/golden/callcounts.json, tests/make_callcounts.py) is run by the host replay with the same number of calls --
    the comp_count stop rule including previous_str carried over from the checkGapReads probes (Figbird.cpp:3919-3927,
    6121-6153).  Speculative probes and candidates past the reference's early exit may add chains, never change one: the reference's
    chains of a (gap, length) must appear among ours in order."""
    import json
    ref = json.load(open(os.path.join(gu.GOLDEN, "callcounts.json")))[os.path.basename(case)][mode]
    dump = os.path.join(case, "oracle_chains_%s.txt" % mode)
    if os.path.exists(dump):
        os.remove(dump)
    fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FB_ORACLE_DUMP": dump}, name="chains")
    mine = _chains(dump)
    want = {}
    for gs, Lg, n in ref:
        want.setdefault((gs, Lg), []).append(n)
    assert want, "empty reference call list"
    def subseq(v, m):      # ours may hold extra chains (speculative probes / candidates past an early exit), in order
        it = iter(m)
        return all(any(x == y for y in it) for x in v)
    bad = [(k, v, mine.get(k, [])) for k, v in want.items() if not subseq(v, mine.get(k, []))]
    assert not bad, "EM chains differ from the reference's (key, reference, ours): %r" % bad[:5]


LIGHT = ("g1", "g2", "g4")      # the host-logic tests below repeat whole runs: small fixtures keep the CPU suite to a few minutes


def _light(case):
    if os.path.basename(case) not in LIGHT:
        pytest.skip("host-logic test runs on the small fixtures only")


def test_sharding_over_two_contexts_is_invariant(case):
    _light(case)
    """N>1 host path on CPU: two engine contexts (FIGBIRD_GPUS=0,1), gaps sharded cost-balanced, same files out."""
    one = fc.run_ours(case, "unmapped", fc.oracle_exe(), name="one")
    two = fc.run_ours(case, "unmapped", fc.oracle_exe(), extra_env={"FIGBIRD_GPUS": "0,1", "FIGBIRD_INFLIGHT": "2", "FIGBIRD_DRAW_THREADS": "3"}, name="two")
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert one[f] == two[f]


def test_tail_speculation_is_invariant(case):
    """Requesting candidates further ahead once a lane's batch is small (the tail of a run) changes what is evaluated
    speculatively, never what is written."""
    if os.path.basename(case) != "g1":
        pytest.skip("one small fixture is enough")
    for mode in ("partial", "unmapped"):
        o = fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FIGBIRD_TAIL_ITEMS": "4096"}, name="tail")
        exp = gu.expected(case, mode)
        for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
            assert o[f] == exp[f], (mode, f)


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_threaded_model_learning_is_identical(case, mode):
    """learnModel cut into blocks on host threads (forced small blocks) gives the same tables and cut-offs as the
    sequential parse, i.e. as the reference (Figbird.cpp:7118-7200)."""
    _light(case)
    model = os.path.join(case, "oracle_model_thr_%s.txt" % mode)
    o = fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FIGBIRD_DUMP_MODEL": model, "FIGBIRD_HOST_THREADS": "5", "FIGBIRD_MODEL_BLOCK": "64"}, name="thr")
    exp = gu.expected(case, mode)
    assert o["gapout.txt"] == exp["gapout.txt"]
    assert gu.model_lines(model) == gu.model_lines(os.path.join(case, "expected", mode, "model.txt"))


def _perturbed_myout(src, dst):
    """myout.sam of a golden case with irregular lines mixed in (whole pairs, so the pairing of pass 2 is kept): soft clips,
    indels, '^' in MD, adjacent MD letters, a pair without tags (inherits the previous MD), extra tags, reordered tags, a
    quality string that looks like a tag, an over-long line, a multi-mapped pair, and a last line without a newline."""
    lines = open(src).read().split("\n")
    lines = [l for l in lines if l]
    out = []
    def edit(l, **kw):
        f = l.split("\t")
        for k, v in kw.items():
            f[{"cigar": 4, "seq": 6, "qual": 7, "md": 8, "ih": 9, "tlen": 5, "flag": 1}[k]] = v
        return f
    for i in range(0, len(lines) - 1, 2):
        a, b = lines[i], lines[i + 1]
        k = (i // 2) % 40
        if k == 3:
            a = "\t".join(edit(a, cigar="20S80M"))
        elif k == 5:
            b = "\t".join(edit(b, cigar="50M2I48M", md="MD:Z:30A67"))
        elif k == 7:
            a = "\t".join(edit(a, cigar="40M3D60M", md="MD:Z:40^ACG60"))
        elif k == 9:
            b = "\t".join(edit(b, md="MD:Z:^AC100"))
        elif k == 11:
            a = "\t".join(edit(a, md="MD:Z:10AC5T82"))
        elif k == 13:
            a, b = "\t".join(a.split("\t")[:8]), "\t".join(b.split("\t")[:8])
        elif k == 15:
            a = a + "\tXS:i:5"
        elif k == 17:
            f = b.split("\t"); b = "\t".join(f[:8] + [f[9], f[8]])
        elif k == 19:
            f = a.split("\t"); f[7] = "MD" + f[7][2:]; a = "\t".join(f)
        elif k == 21:
            f = b.split("\t"); f[7] = "IH:i:" + f[7][5:]; b = "\t".join(f)
        elif k == 23:
            f = a.split("\t"); f[0] = f[0] + "x" * 900; a = "\t".join(f)
        elif k == 25:
            a, b = "\t".join(edit(a, ih="IH:i:2")), "\t".join(edit(b, ih="IH:i:2"))
            out += [a, b]      # the same pair twice: one group for the grouping of pass 2
        elif k == 27:
            a = "\t".join(edit(a, md="MD:Z:0A0C0G97"))
        elif k == 29:
            b = "\t".join(edit(b, seq=b.split("\t")[6][:60] + "N" * 40))
        out += [a, b]
    open(dst, "w").write("\n".join(out))      # (no trailing newline)


@pytest.mark.parametrize("threads", ["1", "5"])
def test_fast_model_parser_equals_generic_on_irregular_input(case, threads):
    """learnModel's in-place parser for regular lines and its strtok-style generic parser (the restatement of
    Figbird.cpp:846-1376) give the same tables and cut-off on a myout.sam with irregular lines mixed in."""
    import shutil
    work = os.path.join(case, "perturbed_%s" % threads)
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.join(work, "Temp"))
    for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
        shutil.copy(os.path.join(case, "unmapped", "Temp", f), os.path.join(work, "Temp", f))
    my = os.path.join(work, "myout.sam")
    _perturbed_myout(os.path.join(case, "unmapped", "myout.sam"), my)
    argv = fc.fillgaps_argv(case, "unmapped", os.path.join(work, "Temp"))
    argv[7] = my
    argv[9] = os.path.join(work, "nogaps") + "/"
    dumps = {}
    for name, env in (("fast", {}), ("generic", {"FIGBIRD_MODEL_GENERIC": "1"})):
        d = os.path.join(work, "model_%s.txt" % name)
        e = dict(os.environ); e.update(env); e.update({"FIGBIRD_DUMP_MODEL": d, "FIGBIRD_HOST_THREADS": threads, "FIGBIRD_MODEL_BLOCK": "64"})
        fc.sh([fc.oracle_exe()] + argv, cwd=work, env=e)
        dumps[name] = open(d).read()
    assert len(dumps["fast"].split("\n")) > 100
    assert dumps["fast"] == dumps["generic"]
