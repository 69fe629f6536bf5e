"""CPU, only where oracle/_ref exists: fresh seeded cases, oracle vs the reference binaries run live."""
import os

import pytest

import fbcase as fc

pytestmark = pytest.mark.skipif(not fc.have_reference(), reason="oracle/_ref not built")


@pytest.mark.parametrize("seed,gen", [
    (41, {"genome": 30000, "gaplist": "20,75", "cov": 25}),
    (42, {"genome": 30000, "gaplist": "3,140", "cov": 35, "readN": 25}),
])
def test_oracle_equals_reference_on_fresh_case(tmp_path, seed, gen):
    g = dict(gen); g["seed"] = seed
    case = fc.make_case(str(tmp_path / "case"), g)
    for mode in ("partial", "unmapped"):
        r = fc.run_reference(case, mode, threads=2)
        o = fc.run_ours(case, mode, fc.oracle_exe())
        for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt"):
            assert r[f] == o[f], "%s %s" % (mode, f)
        assert fc.draw_by_gap(r["draw.txt"]) == fc.draw_by_gap(o["draw.txt"])


@pytest.mark.parametrize("threads", [1, 4, 7, 16])
def test_draw_txt_in_the_reference_worker_order(tmp_path, threads):
    """draw.txt is the concatenation of the workers' files, so its order depends on num_threads and on the reference's dealing of
    small and large gaps to the workers (FillGaps.cpp:456-649): reproduced, byte for byte (partial mode: cheap on the CPU engine)."""
    g = {"genome": 60000, "gaplist": "12,450,45,95,20,520,30,60,25,14,33", "cov": 14, "seed": 43}
    case = fc.make_case(str(tmp_path / "case"), g)
    r = fc.run_reference(case, "partial", threads=threads)
    o = fc.run_ours(case, "partial", fc.oracle_exe(), threads=threads)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert r[f] == o[f], "%s (num_threads %d)" % (f, threads)
    if threads > 1:
        og = fc.run_ours(case, "partial", fc.oracle_exe(), threads=threads, extra_env={"FIGBIRD_DRAW_ORDER": "gap"}, name="gaporder")
        assert fc.draw_by_gap(og["draw.txt"]) == fc.draw_by_gap(r["draw.txt"])
        assert (og["draw.txt"] != r["draw.txt"]) == (threads < 11)      # (one gap per worker is gap order again)


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_draw_txt_is_empty_when_the_first_worker_drew_nothing(tmp_path, threads):
    """mergeFiles (FillGaps.cpp:222-258) builds draw.txt with `out << a.rdbuf() << b.rdbuf()`; an empty `a` puts `out` into the fail
    state and everything after it is lost.  So when the first worker's file is empty and there is more than one worker the
    reference's draw.txt is empty (found on live case L4 of round 1 with 16 workers for 14 gaps); with one worker nothing is merged."""
    g = {"genome": 40000, "gaplist": "15,40,90,25,33", "cov": 14, "seed": 44}
    case = fc.make_case(str(tmp_path / "case"), g)
    open(os.path.join(case, "partial", "Gaps", "partial_gaps_0.sam"), "w").close()      # gap 0 has no reads: nothing is drawn for it
    r = fc.run_reference(case, "partial", threads=threads)
    o = fc.run_ours(case, "partial", fc.oracle_exe(), threads=threads)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert r[f] == o[f], "%s (num_threads %d)" % (f, threads)
    assert (len(r["draw.txt"]) == 0) == (threads >= 5)      # (with 3 workers, worker 0 also has gap 3, which draws)
