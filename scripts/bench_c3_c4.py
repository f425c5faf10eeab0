#!/usr/bin/env python
"""BASELINE configs C3 (bitmap-index build + 3-predicate bitmap AND/OR scan, 500 M rows) and C4 (bitmap equi-join
of a 10 M-row table against a 500 M-row table with COUNT/SUM) on one B200.  Prints one JSON line per config with
kernel times (CUDA events around the library's launches), algorithmic bytes and the fraction of the
measured HBM peak, plus size-independent correctness properties.

    python scripts/bench_c3_c4.py [--rows 500000000] [--reps 5] [--skip-c3] [--skip-c4]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol

N = mbcol._native
SEED = 20260101


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def c3(ctx, rows, reps):
    t = ctx.create_table([(1, 4), (1, 4), (1, 4)], rows)
    t.generate(0, 0, SEED, 1000)     # K: 1 000 distinct
    t.generate(1, 0, SEED, 16)       # G
    t.generate(2, 0, SEED, 4)        # H
    build_ms = {}
    for col, name in ((0, "K"), (1, "G"), (2, "H")):
        t.bitmap_build(col)
        build_ms[name] = ctx.last_kernel_ms
    d = {"K": 1000, "G": 16, "H": 4}
    build = {n: {"ms": ms, "rows_per_s": rows / (ms * 1e-3), "algorithmic_gb": (4 * rows + d[n] * rows / 8) / 1e9,
                 "achieved_gbs": (4 * rows + d[n] * rows / 8) / 1e9 / (ms * 1e-3), "frac_of_measured_peak":
                 (4 * rows + d[n] * rows / 8) / 1e9 / (ms * 1e-3) / peak()} for n, ms in build_ms.items()}
    # properties: every row sets exactly one bit per index; the values are 0..D-1
    assert t.bitmap_values(0).tolist() == list(range(1000)) and t.bitmap_values(2).tolist() == [0, 1, 2, 3]
    terms = [mbcol.Term(N.OP_EQ, ("col", 0), ("int", 17), 0), mbcol.Term(N.OP_EQ, ("col", 1), ("int", 5), 0),
             mbcol.Term(N.OP_EQ, ("col", 2), ("int", 2), 1)]
    times, count = [], None
    for _ in range(reps + 2):
        r = t.bitmap_scan(terms, want=N.WANT_POSITIONS | N.WANT_BITMAP | N.WANT_AGG, aggs=[(0, 0)])
        times.append(ctx.last_kernel_ms)
        count = r.count
        r.close()
    # the same CNF through the row-at-a-time scan must select the same rows (nlj-vs-bmj style cross-check)
    s = t.scan(terms, want=N.WANT_AGG, aggs=[(0, 0)])
    assert s.count == count, (s.count, count)
    s.close()
    # full size against the CPU: the columns of a 1 M-row window regenerated from the counter RNG (oracle.synth_int), the
    # CNF evaluated with numpy, compared with the positions the bitmap scan returned for that window and with the index
    from oracle import oracle as orc
    lo, wn = (rows // 3) // 8192 * 8192, min(1_000_000, rows)
    Kw, Gw, Hw = (orc.synth_int(SEED, c, wn, d, lo) for c, d in ((0, 1000), (1, 16), (2, 4)))
    r = t.bitmap_scan(terms, want=N.WANT_POSITIONS | N.WANT_HOST)
    pos = r.positions()
    a, b = np.searchsorted(pos, lo), np.searchsorted(pos, lo + wn)
    assert np.array_equal(pos[a:b] - lo, np.nonzero(((Kw == 17) | (Gw == 5)) & (Hw == 2))[0])
    r.close()
    bits17 = t.bitmap_get(0, 17)
    got17 = np.nonzero(np.unpackbits(bits17[lo // 64:(lo + wn + 63) // 64].view(np.uint8), bitorder="little")[:wn])[0]
    assert np.array_equal(got17, np.nonzero(Kw == 17)[0])
    ms = statistics.median(times[2:])
    alg = 4 * rows / 8 + 8 * count
    scan = {"ms": ms, "count": count, "selectivity": count / rows, "rows_per_s": rows / (ms * 1e-3), "algorithmic_gb": alg / 1e9,
            "achieved_gbs": alg / 1e9 / (ms * 1e-3), "frac_of_measured_peak": alg / 1e9 / (ms * 1e-3) / peak(),
            "checked_against_cpu": f"positions and the K=17 bitmap on rows [{lo}, {lo + wn}) vs numpy over regenerated columns"}
    t.close()
    return {"config": "C3", "rows": rows, "build": build, "scan": scan}


def c4(ctx, n_r, n_s, reps):
    R = ctx.create_table([(1, 4), (1, 4)], n_r)
    R.generate(0, 3, 0, n_r)                       # key: a permutation of [0, nR)
    R.generate(1, 0, SEED, 1000)                   # v
    S = ctx.create_table([(1, 4), (1, 4), (2, 4)], n_s)
    S.generate(0, 0, SEED + 1, n_r)                # fk uniform over the keys
    S.generate(1, 0, SEED + 2, 1000)               # w
    S.generate(2, 1, SEED + 3)                     # x
    jt = [mbcol.Term(N.OP_EQ, ("col", 0), ("icol", 0), 0)]
    proj = [(1, 0), (1, 1), (2, 1), (2, 2)]
    aggs = [(0, 0), (1, 2), (1, 1), (1, 3)]        # COUNT(*), SUM(S.w), SUM(R.v), SUM(S.x)
    times, vals = [], None
    for _ in range(reps + 2):
        r = mbcol.bitmap_join(R, S, jt, proj, N.WANT_AGG, aggs=aggs)
        times.append(ctx.last_kernel_ms)
        vals = [r.agg(a) for a in range(4)]
        r.close()
    assert vals[0][0] == n_s                        # every fk finds exactly one key
    sw = S.scan([], want=N.WANT_AGG, aggs=[(1, 1), (1, 2)])
    assert vals[1][0] == sw.agg(0)[0] and abs(vals[3][1] - sw.agg(1)[1]) <= 1e-6 * vals[3][1]
    sw.close()
    ms = statistics.median(times[2:])
    alg = 8 * n_r + 12 * n_s
    out = {"config": "C4", "rows_R": n_r, "rows_S": n_s, "ms": ms, "probe_rows_per_s": n_s / (ms * 1e-3), "algorithmic_gb": alg / 1e9,
           "achieved_gbs": alg / 1e9 / (ms * 1e-3), "frac_of_measured_peak": alg / 1e9 / (ms * 1e-3) / peak(),
           "count": vals[0][0], "sum_w": vals[1][0], "sum_v": vals[2][0], "sum_x": vals[3][1]}
    # second run of the config: a 10 % filter on each side
    so = R.scan([mbcol.Term(N.OP_LT, ("col", 1), ("int", 100), 0)], want=N.WANT_BITMAP)
    si = S.scan([mbcol.Term(N.OP_LT, ("col", 1), ("int", 100), 0)], want=N.WANT_BITMAP)
    ft = []
    for _ in range(reps):
        r = mbcol.bitmap_join(R, S, jt, proj, N.WANT_AGG, aggs=aggs, outer_sel=so, inner_sel=si)
        ft.append(ctx.last_kernel_ms)
        fc = r.agg(0)[0]
        r.close()
    out["filtered_10pct"] = {"ms": statistics.median(ft), "count": fc}
    so.close(); si.close(); R.close(); S.close()
    return out


def c4_lowcard(ctx, n_r, n_s, reps):
    """SURVEY 8d, C4's secondary variant: key domain 1 000 on both sides (every key matches ~n_r/1000 outer and ~n_s/1000
    inner rows: ~n_r * n_s / 1000 pairs), aggregates only.  input/BitMapQuery.java:187-305 would walk one inner bitmap per
    outer row; the engine's general equi path reduces both sides per key group.  Checked against the closed form over the
    per-key histograms: COUNT = sum_k nR(k) nS(k), SUM(S.w) = sum_k nR(k) sumW_S(k), SUM(R.v) = sum_k nS(k) sumV_R(k)."""
    R = ctx.create_table([(1, 4), (1, 4)], n_r)
    R.generate(0, 0, SEED + 5, 1000)
    R.generate(1, 0, SEED, 1000)
    S = ctx.create_table([(1, 4), (1, 4), (2, 4)], n_s)
    S.generate(0, 0, SEED + 1, 1000)
    S.generate(1, 0, SEED + 2, 1000)
    S.generate(2, 1, SEED + 3)
    jt = [mbcol.Term(N.OP_EQ, ("col", 0), ("icol", 0), 0)]
    proj = [(1, 0), (1, 1), (2, 1), (2, 2)]
    aggs = [(0, 0), (1, 2), (1, 1), (1, 3)]
    times, vals = [], None
    for _ in range(reps + 2):
        r = mbcol.bitmap_join(R, S, jt, proj, N.WANT_AGG, aggs=aggs)
        times.append(ctx.last_kernel_ms)
        vals = [r.agg(a) for a in range(4)]
        r.close()
    rk, rv = R.read_column(0), R.read_column(1).astype(np.int64)
    sk, sw = S.read_column(0), S.read_column(1).astype(np.int64)
    n_rk = np.bincount(rk, minlength=1000).astype(object)
    n_sk = np.bincount(sk, minlength=1000).astype(object)
    sum_v = np.bincount(rk, weights=None, minlength=1000)          # placeholder shape
    sum_v = np.array([int(x) for x in np.bincount(rk, weights=rv.astype(np.float64), minlength=1000).round()], dtype=object)
    sum_w = np.array([int(x) for x in np.bincount(sk, weights=sw.astype(np.float64), minlength=1000).round()], dtype=object)
    count = int((n_rk * n_sk).sum())
    assert vals[0][0] == count, (vals[0][0], count)
    assert vals[1][0] == int((n_rk * sum_w).sum()) and vals[2][0] == int((n_sk * sum_v).sum())
    ms = statistics.median(times[2:])
    alg = 8 * n_r + 12 * n_s
    R.close(); S.close()
    return {"config": "C4 low-cardinality variant (key domain 1000, aggregates only)", "rows_R": n_r, "rows_S": n_s, "ms": ms,
            "pairs": count, "probe_rows_per_s": n_s / (ms * 1e-3), "algorithmic_gb": alg / 1e9, "achieved_gbs": alg / 1e9 / (ms * 1e-3),
            "frac_of_measured_peak": alg / 1e9 / (ms * 1e-3) / peak(), "checked": "COUNT, SUM(S.w), SUM(R.v) == closed form over the per-key histograms"}


def cpu_baselines(build_rows=2_000_000, n_r=5_000, n_s=250_000):
    """The oracle (literal CPU restatement of the reference, test infrastructure) timed on bounded samples of the same
    workloads on this box's host cores, next to the GPU numbers: a reported baseline, not a target."""
    import time
    from oracle import oracle as orc
    orc.build()
    K = orc.synth_int(SEED, 0, build_rows, 1000)
    t0 = time.perf_counter()
    idx = orc.bitmap_build((1, 4), K)
    tb = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    Rd, Sd = [(1, 4), (1, 4)], [(1, 4), (1, 4), (2, 4)]
    Rc = [orc.synth_perm(n_r, n_r), rng.integers(0, 1000, n_r).astype(np.int32)]
    Sc = [rng.integers(0, n_r, n_s).astype(np.int32), rng.integers(0, 1000, n_s).astype(np.int32), rng.random(n_s).astype(np.float32)]
    jt = [orc.Term(orc.OP_EQ, ("col", 0), ("icol", 0), 0)]
    t0 = time.perf_counter()
    r = orc.bitmap_join(Rd, Rc, Sd, Sc, jt, [(1, 0), (1, 1), (2, 1), (2, 2)], aggs=[(0, 0), (1, 2), (1, 1), (1, 3)])
    tj = time.perf_counter() - t0
    assert len(idx) == 1000 and r["count"] == n_s
    return {"kind": "port", "cores": orc.max_threads(),
            "c3_build_K": {"rows_per_s": build_rows / tb, "sample": f"{build_rows} rows, 1000 distinct values; oracle.bitmap_build (numpy, 1 thread)"},
            "c4_join": {"probe_rows_per_s": n_s / tj, "sample": f"R {n_r} x S {n_s} rows, pair list + tuples + aggregates like BitMapQuery.executeJoin; "
                                                                 f"oracle/mbc_oracle.cpp orc_bitmap_join"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the oracle on bounded samples (host cores)")
    a = ap.parse_args()
    ctx = mbcol.Context(0)
    if not a.skip_c3:
        print(json.dumps(c3(ctx, a.rows, a.reps)), flush=True)
    if not a.skip_c4:
        print(json.dumps(c4(ctx, a.rows // 50, a.rows, a.reps)), flush=True)
        print(json.dumps(c4_lowcard(ctx, a.rows // 50, a.rows, a.reps)), flush=True)
    ctx.close()
    if a.cpu_baseline:
        print(json.dumps({"cpu_baseline": cpu_baselines()}), flush=True)


if __name__ == "__main__":
    main()
