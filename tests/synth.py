"""Seeded synthetic engine-level inputs (model tables + gap batch) for kernel parity tests."""
import numpy as np

from figbird_b200 import capi


def make_model(L=100, mu=200.0, sd=20.0, n_insert=20000, cutoff=6, seed=1, partial=False):
    rng = np.random.default_rng(seed)
    e = np.linspace(0.005, 0.02, L) * (1 + 0.05 * rng.standard_normal(L))
    ins = np.full(L, 1.2e-4) * (1 + 0.1 * rng.random(L))
    dele = np.full(L, 1.1e-4) * (1 + 0.1 * rng.random(L))
    et = rng.random((5, 5)) + 0.2
    et = et / et.sum(axis=1, keepdims=True)
    t = np.arange(n_insert)
    pdf = np.exp(-0.5 * ((t - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi)) + 1e-9
    tmin, tmax = int(mu - 3 * sd), int(mu + 3 * sd)
    if partial:
        tmin, tmax = tmin - L, tmax + L
    return dict(err_pos=e, ins_pos=ins, del_pos=dele, err_type=et, insert_pdf=pdf, insert_min=tmin, insert_max=tmax, prob_cutoff=cutoff)


def make_gap(rng, mode, true_len, og, L=100, n_reads=30, mu=200.0, sd=20.0, gap_start=5000, err=0.01, with_n=False, var_len=False):
    """One gap: truth = left flank + gap sequence + right flank; reads sampled across the gap."""
    F = L
    left = rng.integers(0, 4, F + 400).astype(np.uint8)
    mid = rng.integers(0, 4, true_len).astype(np.uint8)
    right = rng.integers(0, 4, F + 400).astype(np.uint8)
    truth = np.concatenate([left, mid, right])
    gs_t = len(left)                      # truth index of the gap start
    flank = np.concatenate([left[-F:], right[:F]]).astype(np.uint8)
    if with_n:
        flank[rng.integers(0, 2 * F)] = 4
    reads = []
    T = L
    pile_l = np.zeros((T, 4), np.int32); pile_r = np.zeros((T, 4), np.int32)
    for q in range(n_reads):
        ln = L - int(rng.integers(0, 8)) if var_len else L
        if mode == capi.FB_MODE_UNMAPPED:
            # the read lies somewhere over the gap; its mate maps on one side at insert ~ N(mu, sd)
            start = int(rng.integers(gs_t - ln + 5, gs_t + true_len - 5)) if true_len > 10 else int(rng.integers(gs_t - ln + 2, gs_t + max(true_len, 1)))
            codes = truth[start:start + ln].copy()
            flips = rng.random(ln) < err
            codes[flips] = (codes[flips] + rng.integers(1, 4, flips.sum())) % 4
            if with_n and q % 7 == 0:
                codes[int(rng.integers(0, ln))] = 4
            isz = int(round(mu + sd * rng.standard_normal()))
            left_side = bool(rng.integers(0, 2))
            x0 = start - gs_t                      # offset relative to the gap start (truth coordinates)
            if left_side:
                mate = x0 + ln - isz               # insert = x0 - mate + len
                fl = capi.FB_READ_LEFT | (capi.FB_READ_REVERSE if rng.integers(0, 2) else 0)
            else:
                mate = isz - ln + x0 - (true_len - og)   # insert = mate + (Lg-og) + len - x0 at Lg = true_len
                fl = (capi.FB_READ_REVERSE if rng.integers(0, 2) else 0)
            reads.append(dict(codes=codes, mate=mate, flags=fl, jlo=0, jcut=0))
        else:
            # partial read anchored on one flank, overhanging into the gap
            over = int(rng.integers(3, ln - 30))
            if rng.integers(0, 2):
                start = gs_t - (ln - over)
                codes = truth[start:start + ln].copy()
                fl = capi.FB_READ_LEFT
                jlo, jcut = 2, 0
                ci = ln - over - 1
                for t, i in enumerate(range(ci + 1, ln)):
                    if t < T: pile_l[t, codes[i] % 4] += 1
                isz = int(round(mu + sd * rng.standard_normal()))
                mate = (start - gs_t) + ln - isz
            else:
                start = gs_t + true_len - over
                codes = truth[start:start + ln].copy()
                fl = 0
                jlo, jcut = 0, 2
                ci = over
                for u, i in enumerate(range(ci - 1, -1, -1)):
                    if u < T: pile_r[u, codes[i] % 4] += 1
                isz = int(round(mu + sd * rng.standard_normal()))
                mate = isz - ln + (start - gs_t) - (true_len - og)
            flips = rng.random(ln) < err
            codes[flips] = (codes[flips] + rng.integers(1, 4, flips.sum())) % 4
            if q % 5 == 0:
                fl |= capi.FB_READ_NOMATE
            reads.append(dict(codes=codes, mate=mate, flags=fl, jlo=jlo, jcut=jcut))
    return dict(gap_start=gap_start, mode=mode, orig_len=og, flank=flank, pile_left=pile_l, pile_right=pile_r, reads=reads, true_len=true_len)


def compare_results(a, b, rtol_counts=1e-5, cutoff=None, hard=False):
    """a: oracle result, b: device result (dicts of Engine.run).  Returns list of mismatch descriptions.
    cutoff: prob_cutoff of the model.  For EM items the engine guarantees p2max / pos2 only for accepted reads
    (-log10(p) < cutoff); a rejected read must be rejected by both (include/figbird_b200.h, FbItemOut)."""
    bad = []
    for k in ("calls", "comp_count", "flags", "n_reads", "cand_len", "placements"):
        if a[k] != b[k]:
            bad.append("%s: %r vs %r" % (k, a[k], b[k]))
    ns = a["calls"] if a["n_slots"] > 1 else 1
    acc = None
    if cutoff is not None and not hard:
        def accepted(p):
            p = np.asarray(p, dtype=np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                return (p > 0) & (-np.log10(np.where(p > 0, p, 1.0)) < cutoff)
        acc = accepted(a["p2max"][:ns])
        if not np.array_equal(acc, accepted(b["p2max"][:ns])):
            bad.append("accept decisions differ")
    for k in ("pos2", "soft", "hard", "cov"):
        x, y = a[k], b[k]
        if k == "pos2":
            x, y = x[:ns], y[:ns]
            if acc is not None:
                x, y = np.where(acc, x, 0), np.where(acc, y, 0)
        if not np.array_equal(x, y):
            bad.append("%s differs at %s" % (k, np.argwhere(np.asarray(x) != np.asarray(y))[:5].tolist()))
    for k in ("p1max", "p2max"):
        x, y = a[k][:ns], b[k][:ns]
        if k == "p2max" and acc is not None:
            x, y = np.where(acc, x, 0.0), np.where(acc, y, 0.0)
        if k == "p2max":
            # products of exact table entries in the same order: bit-exact
            if not np.array_equal(x, y):
                bad.append("%s not bit-exact (max rel %g)" % (k, np.max(np.abs(x - y) / np.maximum(np.abs(x), 1e-300))))
        else:
            if not np.array_equal(x < 0, y < 0):
                bad.append("p1max validity differs")
            else:
                m = x > 0
                if m.any() and np.max(np.abs(x[m] - y[m]) / x[m]) > 1e-9:
                    bad.append("p1max rel err %g" % np.max(np.abs(x[m] - y[m]) / x[m]))
    if a.get("counts") is not None and b.get("counts") is not None:
        x, y = a["counts"], b["counts"]
        denom = np.maximum(np.abs(x), 1e-300)
        rel = np.abs(x - y) / denom
        rel[(x == 0) & (y == 0)] = 0
        if rel.size and rel.max() > rtol_counts:
            bad.append("counts rel err %g" % rel.max())
    return bad
