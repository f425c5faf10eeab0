"""Parity of the CUDA scan path (K2/K5, through the C ABI) against the CPU oracle.

Bit-exact: qualifying positions (ascending), projected values, reference Tuple bytes, counts, integer
aggregates.  Real SUM: 1e-6 relative (BASELINE.json north_star).  Needs a B200.
"""
import itertools

import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import (C2_AGGS, C2_DESCS, SEED, c2_columns, c2_device_table, c2_terms, check_result, load_table)

pytestmark = pytest.mark.gpu

ALL = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST


def test_synthetic_generator_matches_oracle(ctx, oracle):
    for nrows, base in [(1, 0), (5000, 0), (70001, 123456789), (4096, (1 << 33) + 17)]:
        t = c2_device_table(ctx, nrows, base)
        exp = c2_columns(oracle, nrows, base)
        for c in range(4):
            got = t.read_column(c)
            if c == 2:
                np.testing.assert_array_equal(got.view(np.uint32), exp[c].view(np.uint32))
            else:
                np.testing.assert_array_equal(got, exp[c])
        t.close()
    t = ctx.create_table([(1, 4)], 10007)
    t.generate(0, 3, 0, 10007)
    got = t.read_column(0)
    np.testing.assert_array_equal(got, oracle.synth_perm(10007, 10007))
    assert sorted(got.tolist()) == list(range(10007))
    t.close()


def test_minidata_every_single_column_predicate(ctx, oracle, minidata):
    """BASELINE config C1: every (col, op, literal) through the scan with three projections."""
    names, descs, cols = minidata
    t = load_table(ctx, descs, cols)
    lits = {0: ["Colorado", "South_Dakota", "Zzz"], 1: ["Delaware", "A", "West_Virginia"], 2: [0, 6, 9], 3: [3, -1, 10]}
    n = 0
    for col, op, proj in itertools.product(range(4), range(7), ([0, 1, 2, 3], [2, 0], [3])):
        for lit in lits[col]:
            kind = "int" if descs[col][0] == 1 else "str"
            terms = [oracle.Term(op, ("col", col), (kind, lit), 0)]
            exp = oracle.scan(descs, cols, terms, proj=proj, aggs=[(0, 0), (1, 2), (2, 3), (3, 2)])
            res = t.scan(terms, proj=proj, want=ALL, aggs=[(0, 0), (1, 2), (2, 3), (3, 2)])
            check_result(oracle, res, exp, [descs[c] for c in proj])
            res.close()
            n += 1
    assert n == 4 * 7 * 3 * 3
    t.close()


def test_minidata_no_filter_and_repeated_projection(ctx, oracle, minidata):
    names, descs, cols = minidata
    t = load_table(ctx, descs, cols)
    exp = oracle.scan(descs, cols, [], proj=[3, 3, 0, 2, 0])
    res = t.scan([], proj=[3, 3, 0, 2, 0], want=ALL)
    assert res.count == 500
    check_result(oracle, res, exp, [descs[c] for c in [3, 3, 0, 2, 0]])
    t.close()


def _random_terms(oracle, rng, descs, cols, nconj, maxdis):
    terms = []
    for ci in range(nconj):
        for _ in range(rng.integers(1, maxdis + 1)):
            op = int(rng.integers(0, 7))
            col = int(rng.integers(0, len(descs)))
            t = descs[col][0]
            shape = rng.integers(0, 4)
            if t == 1:
                lit = ("int", int(cols[col][rng.integers(0, len(cols[col]))]) + int(rng.integers(-1, 2)))
            elif t == 2:
                lit = ("real", float(cols[col][rng.integers(0, len(cols[col]))]))
            else:
                row = bytes(cols[col][rng.integers(0, len(cols[col]))]).rstrip(b"\0")
                cut = int(rng.integers(1, len(row) + 1))
                lit = ("str", row[:cut] if shape == 3 else row)
            others = [c for c in range(len(descs)) if descs[c][0] == t and c != col]
            if shape == 1 and others:
                terms.append(oracle.Term(op, ("col", col), ("col", others[rng.integers(0, len(others))]), ci))
            elif shape == 2:
                terms.append(oracle.Term(op, lit, ("col", col), ci))        # literal on the left
            else:
                terms.append(oracle.Term(op, ("col", col), lit, ci))
    return terms


@pytest.mark.parametrize("nrows", [1, 31, 4095, 4096, 4097, 8192, 100003])
def test_random_cnf_ragged_sizes(ctx, oracle, nrows):
    """Random CNFs (AND of ORs, all 7 operators, column-vs-column, literal-on-left, string prefixes) on a
    6-column table; sizes around the tile boundaries."""
    rng = np.random.default_rng(nrows)
    descs = [(1, 4), (1, 4), (2, 4), (2, 4), (0, 16), (0, 16), (0, 5)]
    cols = [rng.integers(-50, 50, nrows).astype(np.int32), rng.integers(-50, 50, nrows).astype(np.int32),
            (rng.integers(0, 64, nrows) / 4).astype(np.float32), (rng.integers(0, 64, nrows) / 4).astype(np.float32)]
    words = [b"ab", b"abc", b"abcd", b"b", b"zz", b"abcdefghijklmnop", b"abcdefghijklmno", b"m"]
    for w in (16, 16, 5):
        pick = rng.integers(0, len(words), nrows)
        cols.append(oracle.pack_strings([words[i][:w].decode() for i in pick], w))
    t = load_table(ctx, descs, cols)
    for q in range(12):
        terms = _random_terms(oracle, rng, descs, cols, int(rng.integers(1, 4)), 3)
        proj = [int(c) for c in rng.integers(0, len(descs), int(rng.integers(1, 6)))]
        aggs = [(0, 0), (1, 0), (1, 2), (2, 1), (3, 3), (2, 2)]
        exp = oracle.scan(descs, cols, terms, proj=proj, aggs=aggs)
        res = t.scan(terms, proj=proj, want=ALL, aggs=aggs)
        check_result(oracle, res, exp, [descs[c] for c in proj])
        res.close()
    t.close()


def test_empty_table(ctx, oracle):
    t = ctx.create_table(C2_DESCS, 0)
    res = t.scan(c2_terms(oracle, 0.1), proj=[0, 3], want=ALL, aggs=C2_AGGS)
    assert res.count == 0 and res.positions().size == 0
    assert res.agg(0) == (0, 0.0, True) and res.agg(1)[0] == 0 and res.agg(3)[2] is False
    t.close()


def test_deleted_rows_are_skipped(ctx, oracle):
    """TupleScan.java:85: rows whose bit is set in markedDeleted never reach the filter."""
    nrows = 50000
    cols = c2_columns(oracle, nrows)
    t = load_table(ctx, C2_DESCS, cols)
    rng = np.random.default_rng(7)
    dele = np.unique(rng.integers(0, nrows, 9000))
    words = oracle.bits_from_positions(dele, nrows)
    t.set_deleted(words)
    for s in (0.01, 0.5, 1.0):
        terms = c2_terms(oracle, s) if s < 1 else []
        exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, deleted_words=words)
        res = t.scan(terms, proj=[0, 1, 2, 3], want=ALL, aggs=C2_AGGS)
        check_result(oracle, res, exp, C2_DESCS)
        assert not np.intersect1d(res.positions(), dele).size
    t.set_deleted(np.zeros(1, dtype=np.uint64))
    assert t.scan([], want=N.WANT_POSITIONS | N.WANT_HOST).count == nrows
    t.close()


@pytest.mark.parametrize("sel", [0.01, 0.1, 0.5])
def test_c2_shape_against_oracle(ctx, oracle, sel):
    """BASELINE config C2 at a size the oracle finishes in seconds (2 M rows), device-generated table."""
    nrows = 2_000_003
    t = c2_device_table(ctx, nrows)
    cols = c2_columns(oracle, nrows)
    terms = c2_terms(oracle, sel)
    exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], aggs=C2_AGGS, nthreads=oracle.max_threads())
    res = t.scan(terms, proj=[0, 1, 2, 3], want=ALL | N.WANT_BITMAP, aggs=C2_AGGS)
    check_result(oracle, res, exp, C2_DESCS)
    assert abs(res.count / nrows - sel) < 0.01
    np.testing.assert_array_equal(oracle.positions_from_bits(res.bitmap(), nrows), exp["positions"])
    t.close()


def test_position_base_offsets_positions(ctx, oracle):
    nrows, base = 30000, (1 << 32) + 5
    t = c2_device_table(ctx, nrows, base)
    cols = c2_columns(oracle, nrows, base)
    exp = oracle.scan(C2_DESCS, cols, c2_terms(oracle, 0.1), proj=[1])
    res = t.scan(c2_terms(oracle, 0.1), proj=[1], want=ALL)
    np.testing.assert_array_equal(res.positions(), exp["positions"] + base)
    np.testing.assert_array_equal(res.column(0), cols[1][exp["positions"]])
    t.close()


def test_scan_host_streaming_matches_resident(ctx, oracle):
    """mbc_scan_host (chunked H2D + scan + D2H) gives the same bytes as the resident scan."""
    nrows = 9_000_011                     # three chunks of 4 Mi rows, ragged tail
    cols = c2_columns(oracle, nrows)
    t = load_table(ctx, C2_DESCS, cols)
    for sel in (0.01, 0.5):
        terms = c2_terms(oracle, sel)
        a = t.scan(terms, proj=[0, 1, 2, 3], want=ALL, aggs=C2_AGGS)
        b = ctx.scan_host(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], want=ALL, aggs=C2_AGGS)
        assert a.count == b.count > 0
        np.testing.assert_array_equal(a.positions(), b.positions())
        for i in range(4):
            np.testing.assert_array_equal(a.column(i), b.column(i))
        np.testing.assert_array_equal(a.tuples(), b.tuples())
        for i in range(len(C2_AGGS)):
            ai, af, av = a.agg(i)
            bi, bf, bv = b.agg(i)
            assert ai == bi and av == bv and abs(af - bf) <= 1e-9 * max(abs(af), 1.0)
        a.close(); b.close()
    t.close()


def _pinned_copy(arr):
    import torch
    buf = torch.empty(arr.nbytes, dtype=torch.uint8, pin_memory=True).numpy()
    out = buf.view(arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def test_scan_host_late_materialisation(ctx, oracle):
    """Pinned host columns + a selective predicate: only the compared columns are uploaded, the survivors' projected /
    aggregated values are read in place from host memory.  Same bytes as the resident scan, fewer bytes over PCIe; a
    dense predicate (sample > 1/5) and pageable buffers take the full-upload path."""
    nrows = 9_000_011
    cols = c2_columns(oracle, nrows)
    pinned = [_pinned_copy(c) for c in cols]
    t = load_table(ctx, C2_DESCS, cols)
    full = sum(w for _, w in C2_DESCS) * nrows
    for sel, host, late in ((0.01, pinned, True), (0.5, pinned, False), (0.01, cols, False)):
        terms = c2_terms(oracle, sel)
        a = t.scan(terms, proj=[3, 1, 0], want=ALL, aggs=C2_AGGS)
        before = ctx.h2d_bytes
        b = ctx.scan_host(C2_DESCS, host, terms, proj=[3, 1, 0], want=ALL, aggs=C2_AGGS)
        moved = ctx.h2d_bytes - before
        assert (moved < 0.5 * full) == late, (sel, moved, full)
        assert a.count == b.count > 0
        np.testing.assert_array_equal(a.positions(), b.positions())
        for i in range(3):
            np.testing.assert_array_equal(a.column(i), b.column(i))
        np.testing.assert_array_equal(a.tuples(), b.tuples())
        for i in range(len(C2_AGGS)):
            ai, af, av = a.agg(i)
            bi, bf, bv = b.agg(i)
            assert ai == bi and av == bv and abs(af - bf) <= 1e-9 * max(abs(af), 1.0)
        a.close(); b.close()
    t.close()


def test_full_size_properties(ctx, oracle):
    """BASELINE config C2 at full size (100 M rows): size-independent properties + an oracle check of a
    1 M-row window regenerated on the host from the counter RNG."""
    nrows = 100_000_000
    t = c2_device_table(ctx, nrows)
    terms = c2_terms(oracle, 0.1)
    want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_BITMAP | N.WANT_HOST
    r1 = t.scan(terms, proj=[0, 1, 2, 3], want=want, aggs=C2_AGGS)
    r2 = t.scan(terms, proj=[0, 1, 2, 3], want=want, aggs=C2_AGGS)
    pos = r1.positions()
    assert r1.count == r2.count == pos.size
    assert abs(r1.count / nrows - 0.1) < 1e-3
    assert np.all(np.diff(pos) > 0) and pos[0] >= 0 and pos[-1] < nrows          # sorted, unique, in range
    np.testing.assert_array_equal(pos, r2.positions())                            # idempotent
    for i in range(len(C2_AGGS)):
        assert r1.agg(i) == r2.agg(i)                                             # reproducible, real SUM included
    bits = r1.bitmap()
    assert int(np.unpackbits(bits.view(np.uint8)).sum()) == r1.count              # popcount(bitmap) == count
    # aggregates are consistent with the projected columns
    assert r1.agg(0)[0] == r1.count
    assert r1.agg(1)[0] == int(r1.column(1).astype(np.int64).sum())
    assert abs(r1.agg(2)[1] - float(r1.column(2).astype(np.float64).sum())) <= 1e-6 * r1.agg(2)[1]
    assert r1.agg(3)[0] == int(r1.column(0).min()) and r1.agg(4)[0] == int(r1.column(0).max())
    # oracle on a window
    lo, n = 73_000_000, 1_000_000
    cols = c2_columns(oracle, n, lo)
    exp = oracle.scan(C2_DESCS, cols, terms, proj=[0, 1, 2, 3], nthreads=oracle.max_threads())
    a, b = np.searchsorted(pos, lo), np.searchsorted(pos, lo + n)
    np.testing.assert_array_equal(pos[a:b] - lo, exp["positions"])
    np.testing.assert_array_equal(r1.column(0)[a:b], cols[0][exp["positions"]])
    np.testing.assert_array_equal(r1.column(3)[a:b], cols[3][exp["positions"]])
    r1.close(); r2.close(); t.close()
