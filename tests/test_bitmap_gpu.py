"""Parity of K3 (bitmap-index build) and K4 (bitmap CNF scan) against the oracle.  Needs a B200."""
import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import check_result, load_table

pytestmark = pytest.mark.gpu

ALL = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_BITMAP | N.WANT_HOST


def _check_index(oracle, t, col, desc, column, deleted_words=None):
    exp = oracle.bitmap_build(desc, column, deleted_words)
    vals = t.bitmap_values(col)
    if desc[0] == 0:
        got_vals = oracle.unpack_strings(vals)
        assert got_vals == sorted(exp.keys(), key=lambda s: s.encode())
    else:
        got_vals = vals.tolist()
        assert got_vals == sorted(exp.keys())
    for v in got_vals:
        np.testing.assert_array_equal(t.bitmap_get(col, v), exp[v], err_msg=f"bitmap of {v!r}")
    return exp


def test_minidata_bitmaps_match_golden_sizes(ctx, oracle, minidata, golden):
    """G13/G14: per-value BitSet.toByteArray().length of the GPU-built bitmaps."""
    names, descs, cols = minidata
    t = load_table(ctx, descs, cols)
    for c in range(4):
        t.bitmap_build(c)
        assert t.bitmap_exists(c)
        _check_index(oracle, t, c, descs[c], cols[c])
    for e in golden:
        if e["kind"] == "index" and e.get("bitmap_bytes"):
            c = names.index(e["cmd"].split()[3])
            vals = t.bitmap_values(c)
            vals = oracle.unpack_strings(vals) if descs[c][0] == 0 else vals.tolist()
            sizes = [oracle.bitset_bytearray_len(t.bitmap_get(c, v)) for v in vals]
            assert sorted(sizes) == sorted(e["bitmap_bytes"])
            if descs[c][0] == 1:
                assert sizes == e["bitmap_bytes"]
    # a value that was never indexed has an empty bitmap (Columnarfile.java:1103-1127)
    assert not t.bitmap_get(2, 12345).any() and not t.bitmap_get(0, "Atlantis").any()
    t.close()


def test_minidata_indexes_query_golden(ctx, oracle, minidata, golden):
    """G10-G12 through mbc_bitmap_scan: rows and order exactly as the reference printed them.
    (Plain CNF semantics here; the duplicate-constraint cache is applied by the ColumnarIndexScan mirror.)"""
    import re
    names, descs, cols = minidata
    t = load_table(ctx, descs, cols)
    for c in range(4):
        t.bitmap_build(c)
    indexes = {c: oracle.bitmap_build(descs[c], cols[c]) for c in range(4)}
    n = 0
    for e in golden:
        if e["kind"] != "indexes_query" or e.get("failed"):
            continue
        m = re.match(r"indexes_query \S+ \S+ \[(.*?)\] (\S+) \d+", e["cmd"])
        targets = [names.index(x) for x in m.group(1).split(",")]
        conj = oracle.parse_cnf(m.group(2), names, descs)
        terms = oracle.cnf_to_terms(conj, descs)
        res = t.bitmap_scan(terms, proj=targets, want=ALL)
        plain = oracle.bitmap_cnf(indexes, names, conj, 500, emulate_duplicate_cache=False)
        np.testing.assert_array_equal(res.positions(), oracle.positions_from_bits(plain, 500))
        np.testing.assert_array_equal(res.bitmap(), plain)
        quirk = oracle.bitmap_cnf(indexes, names, conj, 500, emulate_duplicate_cache=True)
        if np.array_equal(plain, quirk):
            rows = [", ".join(str(v) for v in oracle.decode_tuple(bytes(tp), [descs[c] for c in targets])) for tp in res.tuples()]
            assert rows == e["rows"] and res.count == e["count"]
            n += 1
        res.close()
    assert n >= 4
    t.close()


@pytest.mark.parametrize("nrows,nvalues", [(1, 1), (8191, 3), (8192, 16), (100_003, 1000), (300_000, 3000)])
def test_int_bitmap_build_and_scan(ctx, oracle, nrows, nvalues):
    rng = np.random.default_rng(nrows + nvalues)
    descs = [(1, 4), (1, 4), (1, 4), (2, 4)]
    domain = rng.choice(np.arange(-5000, 5000), size=nvalues, replace=False).astype(np.int32)
    cols = [domain[rng.integers(0, nvalues, nrows)], rng.integers(0, 16, nrows).astype(np.int32),
            rng.integers(0, 4, nrows).astype(np.int32), rng.random(nrows).astype(np.float32)]
    t = load_table(ctx, descs, cols)
    dele = None
    if nrows > 1000:
        dele = oracle.bits_from_positions(np.unique(rng.integers(0, nrows, nrows // 20)), nrows)
        t.set_deleted(dele)
    indexes = {}
    for c in range(3):
        t.bitmap_build(c)
        if nvalues <= 1000 or c > 0:
            indexes[c] = _check_index(oracle, t, c, descs[c], cols[c], dele)
        else:
            indexes[c] = oracle.bitmap_build(descs[c], cols[c], dele)
    names = ["K", "G", "H", "X"]
    k0, k1 = int(domain[0]), int(np.sort(domain)[nvalues // 2])
    queries = [
        f"{{(K,=,{k0})|(G,=,5)}}^{{(H,=,2)}}",                     # BASELINE config C3's shape
        f"{{(K,<,{k1})}}^{{(G,!=,3)|(H,>=,2)}}",
        f"{{(K,>=,{k1})|(K,=,999999)}}",
        f"{{(G,<=,7)}}^{{(G,>,2)}}^{{(H,!=,0)}}",
        "{(K,=,999999)}",                                         # value never indexed -> empty
    ]
    for q in queries:
        conj = oracle.parse_cnf(q, names, descs)
        bits = oracle.bitmap_cnf(indexes, names, conj, nrows, deleted_words=dele, emulate_duplicate_cache=False)
        exp_pos = oracle.positions_from_bits(bits, nrows)
        res = t.bitmap_scan(oracle.cnf_to_terms(conj, descs), proj=[0, 3, 1], want=ALL, aggs=[(0, 0), (1, 1), (1, 3)])
        np.testing.assert_array_equal(res.positions(), exp_pos, err_msg=q)
        np.testing.assert_array_equal(res.bitmap(), bits)
        np.testing.assert_array_equal(res.column(0), cols[0][exp_pos])
        np.testing.assert_array_equal(res.column(1), cols[3][exp_pos])
        assert res.agg(0)[0] == exp_pos.size and res.agg(1)[0] == int(cols[1][exp_pos].astype(np.int64).sum())
        # the bitmap scan and the row-at-a-time scan agree (how the reference's authors cross-checked nlj vs bmj)
        scan = t.scan(oracle.cnf_to_terms(conj, descs), want=N.WANT_POSITIONS | N.WANT_HOST)
        np.testing.assert_array_equal(scan.positions(), exp_pos)
        res.close(); scan.close()
    t.close()


def test_string_bitmap_build(ctx, oracle):
    nrows = 50_001
    rng = np.random.default_rng(3)
    words = ["Alabama", "Alaska", "Arizona", "Arkansas", "California", "Colorado", "a", "ab", "abc", "abcdefghijklmnop"]
    descs = [(0, 16), (0, 7)]
    cols = [oracle.pack_strings([words[i] for i in rng.integers(0, len(words), nrows)], 16),
            oracle.pack_strings([words[i][:7] for i in rng.integers(0, len(words), nrows)], 7)]
    t = load_table(ctx, descs, cols)
    names = ["S", "T"]
    indexes = {}
    for c in range(2):
        t.bitmap_build(c)
        indexes[c] = _check_index(oracle, t, c, descs[c], cols[c])
    for q in ["{(S,=,Colorado)}", "{(S,<=,Arkansas)|(T,=,ab)}", "{(S,>,abc)}^{(T,!=,Alaska)}", "{(S,<,A)}"]:
        conj = oracle.parse_cnf(q, names, descs)
        bits = oracle.bitmap_cnf(indexes, names, conj, nrows, emulate_duplicate_cache=False)
        res = t.bitmap_scan(oracle.cnf_to_terms(conj, descs), proj=[1, 0], want=ALL)
        exp_pos = oracle.positions_from_bits(bits, nrows)
        np.testing.assert_array_equal(res.positions(), exp_pos, err_msg=q)
        np.testing.assert_array_equal(res.column(1), cols[0][exp_pos])
        res.close()
    t.close()


def test_bitmap_scan_without_index_fails_loudly(ctx, oracle):
    t = load_table(ctx, [(1, 4)], [np.arange(100, dtype=np.int32)])
    with pytest.raises(mbcol.MbcError) as ei:
        t.bitmap_scan([oracle.Term(0, ("col", 0), ("int", 5), 0)], want=N.WANT_POSITIONS)
    assert ei.value.status == N.ERR_NOINDEX
    t.close()


@pytest.mark.parametrize("nrows", [8191, 8193, 20_001, 100_003])
def test_few_values_away_from_zero_ragged_tail(ctx, oracle, nrows):
    """<= 32 distinct values that do not include 0 and do not fill their range, no deleted rows, nrows not a multiple of the
    8192-row build chunk: the ragged last chunk sees the column's zero padding, which lies outside [kmin, kmax] and must
    not be looked up in the value -> id table (round 1 read ~16 GB past it)."""
    rng = np.random.default_rng(nrows)
    for values in ([10, 20, 30], [-70000 + 65535, -70000], [5, 6, 9, 1000, 40000]):
        col = rng.choice(np.array(values, dtype=np.int32), nrows).astype(np.int32)
        other = rng.integers(0, 3, nrows).astype(np.int32)
        descs = [(1, 4), (1, 4)]
        t = load_table(ctx, descs, [col, other])
        t.bitmap_build(0)
        _check_index(oracle, t, 0, descs[0], col)
        t.close()
