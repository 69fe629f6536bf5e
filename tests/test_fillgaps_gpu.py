"""-m gpu: the product (`fillgaps` / fb_fillgaps_main on the CUDA engine) against the reference's golden
outputs, plus size-independent properties at larger sizes."""
import os
import re

import pytest

import fbcase as fc
import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=gu.NAMES)
def case(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_gpu_" + request.param)
    return gu.extract(request.param, str(d))


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_product_matches_reference_golden(case, mode):
    o = fc.run_ours(case, mode, fc.product_exe(), extra_env={"FIGBIRD_METRICS": os.path.join(case, "m_%s.json" % mode)})
    exp = gu.expected(case, mode)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert o[f] == exp[f], "%s differs from the reference (%s mode)" % (f, mode)
    import json
    m = json.load(open(os.path.join(case, "m_%s.json" % mode)))
    assert m["engine"] == "cuda-sm100a"
    # the reference's own count of pass-1 placements (counter-instrumented worker, tests/golden/placements.json)
    assert m["ref_placements_p1"] == json.load(open(os.path.join(gu.GOLDEN, "placements.json")))[os.path.basename(case)][mode]


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_device_weights_item_by_item_on_real_inputs(case, mode):
    """The host program on the dual engine (oracle/fb_dual_engine.cpp): every work item of a real run goes to the CUDA library
    and to the CPU oracle; discrete results must be identical, pass-2 products bit-exact, and every per-position base
    weight (countsGap entry) within 1e-5 relative -- BASELINE's bar -- with the files still byte-identical to the reference."""
    import json
    summary = os.path.join(case, "dual_%s.json" % mode)
    env = {"FB_DUAL_DEVICE_LIB": os.path.join(fc.PBUILD, "libfigbird_b200.so"), "FB_DUAL_ORACLE_LIB": os.path.join(fc.OBUILD, "libfb_oracle.so"),
           "FB_DUAL_SUMMARY": summary, "FB_ORACLE_THREADS": "8"}
    o = fc.run_ours(case, mode, os.path.join(fc.OBUILD, "fillgaps_dual"), extra_env=env, name="dual")
    exp = gu.expected(case, mode)
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert o[f] == exp[f], f
    s = json.load(open(summary))
    assert s["items"] > 0 and s["mismatching"] == 0, (s, o["log"][-3000:])
    assert s["worst_weight_rel"] <= 1e-5, s


def test_in_process_entry_point(case):
    from figbird_b200 import capi
    run = os.path.join(case, "run_inproc")
    os.makedirs(os.path.join(run, "Temp"), exist_ok=True)
    import shutil
    for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
        shutil.copy(os.path.join(case, "unmapped", "Temp", f), os.path.join(run, "Temp", f))
    rc = capi.fillgaps(fc.fillgaps_argv(case, "unmapped", os.path.join(run, "Temp")))
    assert rc == 0
    exp = gu.expected(case, "unmapped")
    assert open(os.path.join(run, "Temp", "gapout.txt"), "rb").read() == exp["gapout.txt"]


def test_missing_inputs_exit_1(tmp_path):
    from figbird_b200 import capi
    argv = [str(tmp_path / "nope.fa"), "200", "100", "1", "1", "0", "1", str(tmp_path / "myout.sam"), str(tmp_path) + "/", str(tmp_path) + "/", "30", "100", "0", "0", "200"]
    assert capi.fillgaps(argv) == 1


@pytest.fixture(scope="module")
def big_case(tmp_path_factory):
    if not fc.have_reference():
        pytest.skip("oracle/_ref (reference Preprocess) not available to prepare inputs")
    d = tmp_path_factory.mktemp("big")
    return fc.make_case(str(d / "c1"), {"genome": 1000000, "scaffolds": 4, "gaps": 50, "gapmin": 10, "gapmax": 500, "cov": 30, "sd": 20, "seed": 101, "near": 700, "model-pairs": 100000})


@pytest.mark.parametrize("mode", ["partial", "unmapped"])
def test_properties_at_c1_size(big_case, mode):
    """BASELINE configs[0] size: determinism, invariance to sharding / batching, and output self-consistency."""
    a = fc.run_ours(big_case, mode, fc.product_exe(), name="a")
    b = fc.run_ours(big_case, mode, fc.product_exe(), extra_env={"FIGBIRD_GPUS": "0,0", "FIGBIRD_INFLIGHT": "7"}, name="b")
    for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt"):
        assert a[f] == b[f], f
    # self-consistency of the three files
    draft = open(os.path.join(big_case, "draft.fa")).read().split("\n")
    dlen = sum(len(l) for l in draft if l and not l.startswith(">"))
    delta, n_left = 0, 0
    for line in a["gapout.txt"].decode().split("\n"):
        if not line:
            continue
        t = line.split("\t")
        og, sl, s = int(t[3]), int(t[4]), t[5] if len(t) > 5 else ""
        assert sl == len(s) and re.fullmatch(r"[ACGTN]*", s)
        delta += sl - og
        n_left += s.count("N")
    filled = a["filledContigs.fa"].decode().split("\n")
    flen = sum(len(l) for l in filled if l and not l.startswith(">"))
    assert flen == dlen + delta
    assert a["Ncount.txt"] == (b"0" if n_left == 0 else b"1")
    # most gaps of this synthetic draft get closed or shortened
    assert sum(1 for l in a["gapout.txt"].decode().split("\n") if l and "N" not in l.split("\t")[5]) >= 10


def test_c1_against_live_reference(big_case):
    """Full BASELINE configs[0] differential check (reference runs with numthreads=4 as the config says)."""
    if os.environ.get("FB_SKIP_LIVE_REF"):
        pytest.skip("disabled")
    for mode in ("partial", "unmapped"):
        r = fc.run_reference(big_case, mode, threads=4)
        o = fc.run_ours(big_case, mode, fc.product_exe(), name="live")
        for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt"):
            assert r[f] == o[f], (mode, f)
        assert fc.draw_by_gap(r["draw.txt"]) == fc.draw_by_gap(o["draw.txt"]), mode


def test_c2_full_size_against_reference_golden(tmp_path_factory):
    """BASELINE configs[1] at full size (500 gaps): gapout.txt of both modes byte-identical to what the reference
    FillGaps + Figbird (-O2 worker, 8 threads) wrote for the same seeded inputs (tests/golden/c2_*.gz, produced
    by running oracle/_ref in the dev container; inputs are regenerated here from the same seed)."""
    if not fc.have_reference():
        pytest.skip("oracle/_ref (reference Preprocess) not available to prepare inputs")
    import gzip
    import hashlib
    import bench
    d = tmp_path_factory.mktemp("c2")
    case = bench.prepare_case(str(d / "c2"), bench.WORKLOADS["c2n"], 102)
    md5s = open(os.path.join(gu.GOLDEN, "c2_filled.md5")).read().split()
    for i, mode in enumerate(("partial", "unmapped")):
        o = fc.run_ours(case, mode, fc.product_exe(), threads=8)
        exp = gzip.open(os.path.join(gu.GOLDEN, "c2_gapout_%s.txt.gz" % mode)).read()
        assert o["gapout.txt"] == exp, "c2 %s gapout differs from the reference" % mode
        assert hashlib.md5(o["filledContigs.fa"]).hexdigest() == md5s[i]


def test_pipeline_our_preprocess_then_our_fillgaps(tmp_path):
    """The two stages chained without any reference program: bowtie2-style SAM -> fb_preprocess_main -> fb_fillgaps_main, against
    what the reference Preprocess + reference FillGaps made of the same SAM (tests/golden/pp1.tar.gz, tests/make_pp_golden.py)."""
    import hashlib
    import json
    import shutil
    import tarfile
    from figbird_b200 import capi
    with tarfile.open(os.path.join(gu.GOLDEN, "pp1.tar.gz")) as t:
        t.extractall(str(tmp_path), filter="data")
    case = str(tmp_path / "pp1")
    exp = json.load(open(os.path.join(case, "expected.json")))
    if "fillgaps_partial" not in exp:
        pytest.skip("fixture without FillGaps expectations")
    draft = os.path.join(case, "draft.fa")
    gaps = os.path.join(case, "Gaps") + "/"        # ONE Gaps/ directory for both passes, as RunFigbird.sh keeps it
    os.makedirs(gaps)
    for mode, flag, sam in (("partial", "1", "result1.sam"), ("unmapped", "2", "result2.sam")):
        d = os.path.join(case, mode); tmp = os.path.join(d, "Temp") + "/"
        os.makedirs(tmp)
        assert capi.tool("preprocess", [draft, str(exp["x"][mode]), flag, os.path.join(case, sam), os.path.join(d, "myout.sam"), draft, "r1.fq", "r2.fq", gaps, tmp, "1", "0", "0"]) == 0
        pf, um = ("1", "0") if mode == "partial" else ("0", "1")
        argv = [draft, str(exp["x"][mode]), str(exp["readlen"]), "1", pf, um, "4", os.path.join(d, "myout.sam"), tmp, gaps, "30", str(exp["readlen"]), "0", "0", str(exp["insert"])]
        assert capi.fillgaps(argv) == 0
        out = fc.read_outputs(tmp)
        for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt"):
            assert hashlib.md5(out[f]).hexdigest() == exp["fillgaps_" + mode][f], "%s %s" % (mode, f)
        assert hashlib.md5(repr(sorted(fc.draw_by_gap(out["draw.txt"]).items())).encode()).hexdigest() == exp["fillgaps_" + mode]["draw.txt (per gap, sorted)"], mode
    shutil.rmtree(case, ignore_errors=True)
