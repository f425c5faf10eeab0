"""Every scan engine of libmbcol.so against the CPU oracle, through the C ABI.

`mbc_scan` has the two-pass engine (filter -> offsets -> write: write_kernel writes the groups of 8 tiles with few survivors
whole, one CTA each, and gathers the survivors of fuller groups tile by tile; the fullest groups -- above 1/3 -- go through
the TMA-staged write_staged_kernel unless MBC_WRITE_STAGED=0 or a projected row is too wide for its ring) and the opt-in single-residency
engine (mbc_scan_fused.cuh, MBC_SCAN_PATH=fused: count warps ahead, offsets from published tile counts, compaction out of
shared memory).  Every case here runs under all of them and is compared with the oracle: bit-exact positions / values /
Tuple bytes / integer aggregates, real SUM 1e-6 relative.
Needs a B200.
"""
import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import C2_AGGS, C2_DESCS, c2_columns, c2_device_table, c2_terms, check_result, load_table

pytestmark = pytest.mark.gpu

ALL = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST
ENGINES = ["twopass", "staged2", "gather", "fused"]


@pytest.fixture(params=ENGINES)
def engine(request, monkeypatch):
    """twopass: the default engine (filter -> offsets -> write_kernel for the groups up to 1/3 full + write_staged_kernel for
    the fuller ones); staged2: the same with the shallowest ring (two stages) and every group above 12.5 % staged; gather:
    write_staged_kernel off, groups above 12.5 % gathered tile by tile (the round-1 write pass); fused: the single-residency
    kernel (opt-in)."""
    e = request.param
    monkeypatch.setenv("MBC_SCAN_PATH", "fused" if e == "fused" else "twopass")
    monkeypatch.setenv("MBC_WRITE_STAGED", "0" if e == "gather" else "1")
    if e == "staged2":
        monkeypatch.setenv("MBC_STAGED_STAGES", "2")
        monkeypatch.setenv("MBC_STAGED_MIN_PCT", "0")
    return e


@pytest.mark.parametrize("sel", [0.0005, 0.01, 0.05, 0.1, 0.5, 0.9, 1.0])
def test_c2_selectivities(ctx, oracle, engine, sel):
    """C2 shape, 2 M rows (977 fused tiles = 6.6 waves on 148 SMs, ragged last tile): all-sparse, mixed and all-dense tiles."""
    nrows = 2_000_003
    t = c2_device_table(ctx, nrows)
    cols = c2_columns(oracle, nrows)
    terms = c2_terms(oracle, sel) if sel < 1 else []
    exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
    res = t.scan(terms, proj=[0, 1, 2, 3], want=ALL | N.WANT_BITMAP, aggs=C2_AGGS)
    check_result(oracle, res, exp, C2_DESCS)
    np.testing.assert_array_equal(oracle.positions_from_bits(res.bitmap(), nrows), exp["positions"])
    res.close(); t.close()


@pytest.mark.parametrize("dense_min", [1, 64, 1024])
def test_dense_threshold_extremes(ctx, oracle, engine, monkeypatch, dense_min):
    """Every tile bulk-copied whole (1), the default mix, every tile below half full gathered in batches (1024)."""
    monkeypatch.setenv("MBC_FUSED_DENSE_MIN", str(dense_min))
    nrows = 700_001
    t = c2_device_table(ctx, nrows)
    cols = c2_columns(oracle, nrows)
    for sel in (0.003, 0.2, 0.45):
        terms = c2_terms(oracle, sel)
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[3, 0, 2], aggs=C2_AGGS, nthreads=oracle.max_threads())
        res = t.scan(terms, proj=[3, 0, 2], want=ALL, aggs=C2_AGGS)
        check_result(oracle, res, exp, [C2_DESCS[c] for c in (3, 0, 2)])
        res.close()
    t.close()


def test_clustered_table_mixes_dense_sparse_and_empty_tiles(ctx, oracle, engine):
    """Sorted predicate column: a dense prefix, a ragged boundary tile, then runs of sparse and empty tiles."""
    nrows = 1_500_017
    rng = np.random.default_rng(5)
    i1 = np.sort(rng.integers(0, 1 << 20, nrows)).astype(np.int32)
    i2 = rng.integers(-1000, 1000, nrows).astype(np.int32)
    r = (rng.integers(0, 4000, nrows) / 4).astype(np.float32)
    s = oracle.pack_strings(["k%07d" % (x % 9973) for x in range(nrows)], 16)
    cols = [i1, i2, r, s]
    t = load_table(ctx, C2_DESCS, cols)
    cut = int(i1[nrows // 3])
    cases = [
        [oracle.Term(oracle.OP_LT, ("col", 0), ("int", cut), 0)],                                             # dense prefix, empty rest
        [oracle.Term(oracle.OP_LT, ("col", 0), ("int", cut), 0), oracle.Term(oracle.OP_LT, ("col", 1), ("int", -990), 0)],   # dense prefix OR sparse everywhere
        [oracle.Term(oracle.OP_GE, ("col", 0), ("int", cut), 0), oracle.Term(oracle.OP_EQ, ("col", 1), ("int", 7), 1)],      # empty prefix, very sparse rest
        [oracle.Term(oracle.OP_EQ, ("col", 1), ("int", 123456), 0)],                                          # nothing qualifies
    ]
    for terms in cases:
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
        res = t.scan(terms, proj=[0, 1, 2, 3], want=ALL, aggs=C2_AGGS)
        check_result(oracle, res, exp, C2_DESCS)
        res.close()
    t.close()


def test_more_waves_than_mask_slots(ctx, oracle, engine):
    """6 M rows = 2 930 fused tiles = 20 waves on 148 SMs: every ring of the fused kernel wraps at least once."""
    nrows = 6_000_000
    t = c2_device_table(ctx, nrows)
    cols = c2_columns(oracle, nrows)
    for sel in (0.02, 0.5):
        terms = c2_terms(oracle, sel)
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
        res = t.scan(terms, proj=[0, 1, 2, 3], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_HOST, aggs=C2_AGGS)
        check_result(oracle, res, exp, C2_DESCS, want_tuples=False)
        res.close()
    t.close()


@pytest.mark.parametrize("nrows", [1, 2047, 2048, 2049, 4096, 10_000, 303_105])
def test_shapes_of_what_is_asked_for(ctx, oracle, engine, nrows):
    """Positions only, aggregates only, COUNT only, repeated projections, odd string widths (strides 8 and 48)."""
    rng = np.random.default_rng(nrows)
    descs = [(1, 4), (2, 4), (0, 5), (0, 40), (1, 4)]
    cols = [rng.integers(-20, 20, nrows).astype(np.int32), (rng.integers(0, 100, nrows) / 8).astype(np.float32),
            oracle.pack_strings(["w%d" % (x % 50) for x in rng.integers(0, 1000, nrows)], 5),
            oracle.pack_strings(["long-value-%d-%s" % (x, "x" * (x % 20)) for x in rng.integers(0, 1000, nrows)], 40),
            rng.integers(0, 1 << 30, nrows).astype(np.int32)]
    t = load_table(ctx, descs, cols)
    terms = [oracle.Term(oracle.OP_GE, ("col", 0), ("int", 3), 0), oracle.Term(oracle.OP_LT, ("col", 1), ("real", 3.0), 0),
             oracle.Term(oracle.OP_NE, ("col", 2), ("str", "w7"), 1)]
    aggs = [(1, 4), (2, 0), (3, 1), (0, 0), (1, 1)]
    for proj, want, ag in (([], N.WANT_POSITIONS | N.WANT_HOST, []),
                           ([], N.WANT_AGG | N.WANT_HOST, aggs),
                           ([], N.WANT_AGG | N.WANT_HOST, [(0, 0)]),
                           ([3, 2, 3, 0], ALL, aggs),
                           ([4], N.WANT_COLUMNS | N.WANT_HOST, []),
                           ([1, 2], ALL, [(0, 0)])):
        exp = oracle.scan(descs, cols, terms, proj=proj, aggs=ag)
        res = t.scan(terms, proj=proj, want=want, aggs=ag)
        assert res.count == exp["count"]
        if want & N.WANT_POSITIONS:
            np.testing.assert_array_equal(res.positions(), exp["positions"])
        if want == ALL:
            check_result(oracle, res, exp, [descs[c] for c in proj])
        elif want & N.WANT_COLUMNS:
            np.testing.assert_array_equal(res.column(0), cols[proj[0]][exp["positions"]])
        if want & N.WANT_AGG:
            for a, (ei, ef, ev) in enumerate(exp["aggs"]):
                gi, gf, gv = res.agg(a)
                assert gv == ev
                if ev:
                    assert gi == ei if float(ei) == ef else abs(gf - ef) <= 1e-6 * max(abs(ef), 1e-30)
        res.close()
    t.close()


def test_deleted_rows_and_position_base(ctx, oracle, engine):
    nrows, base = 400_000, (1 << 33) + 11
    cols = c2_columns(oracle, nrows, base)
    t = load_table(ctx, C2_DESCS, cols, base)
    rng = np.random.default_rng(3)
    dele = np.unique(rng.integers(0, nrows, 60_000))
    words = oracle.bits_from_positions(dele, nrows)
    t.set_deleted(words)
    for sel in (0.01, 0.6):
        terms = c2_terms(oracle, sel)
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[1, 3], aggs=C2_AGGS, deleted_words=words)
        res = t.scan(terms, proj=[1, 3], want=ALL, aggs=C2_AGGS)
        assert res.count == exp["count"]
        np.testing.assert_array_equal(res.positions(), exp["positions"] + base)
        np.testing.assert_array_equal(res.column(0), cols[1][exp["positions"]])
        np.testing.assert_array_equal(res.column(1), cols[3][exp["positions"]])
        np.testing.assert_array_equal(res.tuples(), exp["tuples"])
    t.close()


def test_scan_host_against_the_oracle(ctx, oracle, engine):
    """mbc_scan_host (chunked H2D + scan + D2H; every chunk appends at the running count) directly against the oracle."""
    nrows = 9_000_011                     # three chunks of 4 Mi rows, ragged tail
    cols = c2_columns(oracle, nrows)
    for sel in (0.01, 0.5):
        terms = c2_terms(oracle, sel)
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
        res = ctx.scan_host(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], want=ALL, aggs=C2_AGGS)
        check_result(oracle, res, exp, C2_DESCS)
        res.close()


def _pinned_copy(arr):
    import torch
    buf = torch.empty(arr.nbytes, dtype=torch.uint8, pin_memory=True).numpy()
    out = buf.view(arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def test_scan_host_late_materialisation_against_the_oracle(ctx, oracle, engine):
    """Pinned host columns + a selective predicate (survivors' values read in place over PCIe) against the oracle."""
    nrows = 9_000_011
    cols = c2_columns(oracle, nrows)
    pinned = [_pinned_copy(c) for c in cols]
    terms = c2_terms(oracle, 0.01)
    exp = oracle.scan(C2_DESCS, cols, terms, proj=[3, 1, 0], aggs=C2_AGGS, nthreads=oracle.max_threads())
    before = ctx.h2d_bytes
    res = ctx.scan_host(C2_DESCS, pinned, terms, proj=[3, 1, 0], want=ALL, aggs=C2_AGGS)
    assert ctx.h2d_bytes - before < 0.5 * 28 * nrows
    check_result(oracle, res, exp, [C2_DESCS[c] for c in (3, 1, 0)])
    res.close()


def test_repeat_200_times_is_bit_identical(ctx, oracle, engine):
    """The only race detector this pool offers (compute-sanitizer is closed): the three C2 scans, 200 times on one 2 M-row
    table, must give the same count, positions, bitmap, columns and aggregates (real SUM included) every time."""
    import hashlib
    nrows = 2_000_003
    t = c2_device_table(ctx, nrows)
    want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_BITMAP | N.WANT_HOST
    first = {}
    for it in range(200):
        for sel in (0.01, 0.1, 0.5):
            res = t.scan(c2_terms(oracle, sel), proj=[0, 1, 2, 3], want=want, aggs=C2_AGGS)
            h = hashlib.sha256()
            h.update(res.positions().tobytes())
            h.update(res.bitmap().tobytes())
            for c in range(4):
                h.update(np.ascontiguousarray(res.column(c)).tobytes())
            sig = (res.count, h.hexdigest(), tuple(res.agg(a) for a in range(len(C2_AGGS))))
            res.close()
            if it == 0:
                first[sel] = sig
            else:
                assert sig == first[sel], (it, sel)
    cols = c2_columns(oracle, nrows)
    for sel in (0.01, 0.1, 0.5):
        exp = oracle.scan(C2_DESCS, cols, c2_terms(oracle, sel), proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
        assert first[sel][0] == exp["count"]
    t.close()
