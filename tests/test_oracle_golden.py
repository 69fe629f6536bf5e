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
    o = fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FB_ORACLE_DUMP": dump, "FIGBIRD_DUMP_MODEL": model}, name="oracle")
    exp = gu.expected(case, mode)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert o[f] == exp[f], "%s differs from the reference (%s mode)" % (f, mode)
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


def test_sharding_over_two_contexts_is_invariant(case):
    """N>1 host path on CPU: two engine contexts (FIGBIRD_GPUS=0,1), gaps sharded cost-balanced, same files out."""
    one = fc.run_ours(case, "unmapped", fc.oracle_exe(), name="one")
    two = fc.run_ours(case, "unmapped", fc.oracle_exe(), extra_env={"FIGBIRD_GPUS": "0,1", "FIGBIRD_INFLIGHT": "2"}, name="two")
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert one[f] == two[f]


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_threaded_model_learning_is_identical(case, mode):
    """learnModel cut into blocks on host threads (forced small blocks) gives the same tables and cut-offs as the
    sequential parse, i.e. as the reference (Figbird.cpp:7118-7200)."""
    model = os.path.join(case, "oracle_model_thr_%s.txt" % mode)
    o = fc.run_ours(case, mode, fc.oracle_exe(), extra_env={"FIGBIRD_DUMP_MODEL": model, "FIGBIRD_HOST_THREADS": "5", "FIGBIRD_MODEL_BLOCK": "64"}, name="thr")
    exp = gu.expected(case, mode)
    assert o["gapout.txt"] == exp["gapout.txt"]
    assert gu.model_lines(model) == gu.model_lines(os.path.join(case, "expected", mode, "model.txt"))
