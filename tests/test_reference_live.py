"""CPU, only where oracle/_ref exists: fresh seeded cases, oracle vs the reference binaries run live."""
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
