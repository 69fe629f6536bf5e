"""-m gpu: the CUDA engine against the CPU oracle engine, item by item, through the C ABI.

Bar (BASELINE.json north_star): discrete outputs identical (positions, consensus strings, coverage,
round counts, flags), pass-2 products bit-exact, per-position base weights (countsGap) within 1e-5 relative.
"""
import os

import numpy as np
import pytest

import synth
from figbird_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libfb_oracle.so")
pytestmark = pytest.mark.gpu


def _pair():
    dev = capi.Engine(0)                      # raises when the CUDA library / GPU is missing: no fallback
    assert dev.name() == "cuda-sm100a"
    ora = capi.Engine(0, lib_path=ORACLE_LIB)
    assert ora.name() == "oracle-cpu"
    return dev, ora


def _run_both(model, gaps, items):
    dev, ora = _pair()
    try:
        for e in (dev, ora):
            e.upload_model(**model)
            e.upload_batch(gaps, None)
        a = ora.run(items)
        b = dev.run(items)
    finally:
        dev.close(); ora.close()
    return a, b


def _assert_same(a, b, items, model=None):
    for i, (x, y) in enumerate(zip(a, b)):
        bad = synth.compare_results(x, y, cutoff=None if model is None else model["prob_cutoff"], hard=items[i].get("kind", 0) == capi.FB_ITEM_HARD)
        assert not bad, "item %d %r: %s" % (i, {k: v for k, v in items[i].items() if k not in ("counts_in", "string_in")}, "; ".join(bad))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_unmapped_em_items(seed):
    rng = np.random.default_rng(seed)
    model = synth.make_model(seed=seed)
    gaps = [synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 60, 55, n_reads=30),
            synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 20, 24, n_reads=25, with_n=True),
            synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 150, 140, n_reads=60, var_len=True),
            synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 8, 10, n_reads=12)]
    items = []
    for gi, g in enumerate(gaps):
        for Lg in sorted({0, 1, g["true_len"] - 3, g["true_len"], g["true_len"] + 9, 3 * g["orig_len"]}):
            if Lg < 0:
                continue
            items.append(dict(gap=gi, cand_len=Lg, max_rounds=200, flags=capi.FB_FLAG_EXTRA_PASS | capi.FB_FLAG_WANT_COUNTS))
            items.append(dict(gap=gi, cand_len=Lg, max_rounds=200, flags=capi.FB_FLAG_WANT_COUNTS))
    a, b = _run_both(model, gaps, items)
    _assert_same(a, b, items, model)
    assert any(r["calls"] > 3 for r in b)


@pytest.mark.parametrize("seed", [4, 5])
def test_partial_em_items(seed):
    rng = np.random.default_rng(seed)
    model = synth.make_model(seed=seed, partial=True)
    gaps = [synth.make_gap(rng, capi.FB_MODE_PARTIAL, 50, 50, n_reads=40),
            synth.make_gap(rng, capi.FB_MODE_PARTIAL, 12, 15, n_reads=20),
            synth.make_gap(rng, capi.FB_MODE_PARTIAL, 230, 210, n_reads=80, var_len=True)]
    items = []
    fl = capi.FB_FLAG_RECORD_ALL | capi.FB_FLAG_NO_COMP_STOP | capi.FB_FLAG_WANT_COUNTS
    for gi, g in enumerate(gaps):
        for Lg in sorted({0, 2, g["true_len"] - 1, g["true_len"], g["true_len"] + 17, 300}):
            items.append(dict(gap=gi, cand_len=Lg, max_rounds=3, flags=fl))
    a, b = _run_both(model, gaps, items)
    _assert_same(a, b, items, model)
    assert all(r["calls"] == 3 and r["n_slots"] == 3 for r in b)


def test_hard_items_and_resume():
    rng = np.random.default_rng(9)
    model = synth.make_model(seed=9)
    gaps = [synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 70, 66, n_reads=35),
            synth.make_gap(rng, capi.FB_MODE_PARTIAL, 40, 44, n_reads=30)]
    dev, ora = _pair()
    try:
        for e in (dev, ora):
            e.upload_model(**model); e.upload_batch(gaps, None)
        # one round, then resume from the returned counts (the host-driven large-gap path)
        first = [dict(gap=0, cand_len=70, max_rounds=1, flags=capi.FB_FLAG_WANT_COUNTS)]
        a0, b0 = ora.run(first), dev.run(first)
        _assert_same(a0, b0, first, model)
        nxt = [dict(gap=0, cand_len=70, max_rounds=1, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_RESUME, comp_count_in=a0[0]["comp_count"],
                    counts_in=a0[0]["counts"], string_in=a0[0]["hard"])]
        a1, b1 = ora.run(nxt), dev.run(nxt)
        _assert_same(a1, b1, nxt, model)
        # two single rounds == one two-round item
        two = [dict(gap=0, cand_len=70, max_rounds=2, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_NO_COMP_STOP)]
        b2 = dev.run(two)
        assert np.array_equal(b2[0]["hard"], b1[0]["hard"]) and np.array_equal(b2[0]["pos2"], b1[0]["pos2"])
        # hard placement on a given string, both finalize semantics
        hard = [dict(kind=capi.FB_ITEM_HARD, gap=0, cand_len=70, string_in=a1[0]["soft"]),
                dict(kind=capi.FB_ITEM_HARD, gap=1, cand_len=40, string_in=rng.integers(0, 5, 40).astype(np.uint8), flags=capi.FB_FLAG_FINALIZE_REF),
                dict(kind=capi.FB_ITEM_HARD, gap=1, cand_len=52, string_in=rng.integers(0, 4, 52).astype(np.uint8))]
        a3, b3 = ora.run(hard), dev.run(hard)
        _assert_same(a3, b3, hard, model)
    finally:
        dev.close(); ora.close()


def test_large_candidate_uses_global_tables():
    """Lg large enough that the row tables no longer fit in shared memory."""
    rng = np.random.default_rng(11)
    model = synth.make_model(seed=11)
    gaps = [synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 1500, 1400, n_reads=120)]
    items = [dict(gap=0, cand_len=1500, max_rounds=2, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_NO_COMP_STOP),
             dict(gap=0, cand_len=2600, max_rounds=1, flags=capi.FB_FLAG_WANT_COUNTS)]
    a, b = _run_both(model, gaps, items)
    _assert_same(a, b, items, model)


def test_deterministic_across_runs():
    rng = np.random.default_rng(12)
    model = synth.make_model(seed=12)
    gaps = [synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 90, 80, n_reads=50)]
    items = [dict(gap=0, cand_len=L, max_rounds=200, flags=capi.FB_FLAG_EXTRA_PASS | capi.FB_FLAG_WANT_COUNTS) for L in (80, 90, 100)]
    dev = capi.Engine(0)
    try:
        dev.upload_model(**model); dev.upload_batch(gaps, None)
        r1 = dev.run(items); r2 = dev.run(list(reversed(items)))
    finally:
        dev.close()
    for x, y in zip(r1, reversed(r2)):
        assert np.array_equal(x["counts"], y["counts"]) and np.array_equal(x["p1max"], y["p1max"])


def test_partial_subnormal_weights():
    """Partial-mode weights pow(10, ln p) underflow gradually; rows whose only votes are subnormal still decide the
    reference's consensus, so the device must round them like glibc (regression: found at C2 scale by the dual engine)."""
    rng = np.random.default_rng(21)
    model = synth.make_model(seed=21, partial=True)
    gaps = []
    for _ in range(6):
        g = synth.make_gap(rng, capi.FB_MODE_PARTIAL, 180, 170, n_reads=40)
        # make half of the reads unrelated to the locus: their products fall to 1e-130..1e-160, weights to ~1e-320
        for r in g["reads"][::2]:
            r["codes"] = rng.integers(0, 4, len(r["codes"])).astype(np.uint8)
        gaps.append(g)
    fl = capi.FB_FLAG_RECORD_ALL | capi.FB_FLAG_NO_COMP_STOP | capi.FB_FLAG_WANT_COUNTS
    items = [dict(gap=gi, cand_len=Lg, max_rounds=3, flags=fl) for gi in range(len(gaps)) for Lg in (120, 150, 170, 180, 200, 260)]
    a, b = _run_both(model, gaps, items)
    sub = sum(int(((r["counts"] > 0) & (r["counts"] < 2.3e-308)).sum()) for r in a)
    assert sub > 0, "test does not reach the subnormal range"
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x["soft"], y["soft"]), "item %d soft consensus" % i
        assert not [m for m in synth.compare_results(x, y, cutoff=model["prob_cutoff"]) if not m.startswith("counts")]
        m = (x["counts"] > 0) & (x["counts"] < 2.3e-308)
        assert np.array_equal(x["counts"][m], y["counts"][m]), "subnormal weights must match bit for bit"


def test_many_reads_take_the_chunked_path():
    """Read sets whose weight rows do not fit the chunk region at once (the cap of the reference is 3000 reads per gap,
    Figbird.cpp:114-115): reads are staged and gathered chunk by chunk; results must not depend on the chunking."""
    rng = np.random.default_rng(31)
    model_u = synth.make_model(seed=31)
    gaps = [synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 120, 110, n_reads=700),
            synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 40, 44, n_reads=3000)]
    items = [dict(gap=0, cand_len=120, max_rounds=3, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_NO_COMP_STOP),
             dict(gap=0, cand_len=33, max_rounds=2, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_NO_COMP_STOP | capi.FB_FLAG_EXTRA_PASS),
             dict(gap=1, cand_len=40, max_rounds=2, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_NO_COMP_STOP)]
    a, b = _run_both(model_u, gaps, items)
    _assert_same(a, b, items, model_u)
    model_p = synth.make_model(seed=32, partial=True)
    gaps = [synth.make_gap(rng, capi.FB_MODE_PARTIAL, 90, 80, n_reads=900)]
    fl = capi.FB_FLAG_RECORD_ALL | capi.FB_FLAG_NO_COMP_STOP | capi.FB_FLAG_WANT_COUNTS
    items = [dict(gap=0, cand_len=L, max_rounds=3, flags=fl) for L in (0, 60, 90, 250)]
    a, b = _run_both(model_p, gaps, items)
    _assert_same(a, b, items, model_p)


def test_empty_and_short_inputs():
    """A gap without reads, reads much shorter than the model's read length, and candidate lengths around the read length."""
    rng = np.random.default_rng(33)
    model = synth.make_model(seed=33)
    g_empty = synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 30, 30, n_reads=0)
    g_short = synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 50, 45, L=36, n_reads=40)
    g_edge = synth.make_gap(rng, capi.FB_MODE_UNMAPPED, 100, 100, n_reads=30)
    gaps = [g_empty, g_short, g_edge]
    items = [dict(gap=0, cand_len=L, max_rounds=3, flags=capi.FB_FLAG_WANT_COUNTS) for L in (0, 30)]
    items += [dict(gap=1, cand_len=L, max_rounds=200, flags=capi.FB_FLAG_WANT_COUNTS | capi.FB_FLAG_EXTRA_PASS) for L in (1, 35, 36, 37, 50, 140)]
    items += [dict(gap=2, cand_len=L, max_rounds=200, flags=capi.FB_FLAG_WANT_COUNTS) for L in (31, 32, 33, 63, 64, 65, 99, 100, 101, 127, 128, 129)]
    items += [dict(kind=capi.FB_ITEM_HARD, gap=0, cand_len=30, string_in=rng.integers(0, 4, 30).astype(np.uint8))]
    a, b = _run_both(model, gaps, items)
    _assert_same(a, b, items, model)
