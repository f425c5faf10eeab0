"""Parity of the GPU sort (mbc_sort, SURVEY 8f rank 4: input/ColumnarSort.java) against the oracle and the reference's
golden `sort` runs: positions in key order, ties by ascending position, projected fields and Tuple bytes in that order."""
import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import load_table

pytestmark = pytest.mark.gpu
ALL = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_HOST


def _table(oracle, n, seed):
    rng = np.random.default_rng(seed)
    words = ["", "a", "ab", "abc", "b", "Zulu", "alpha", "alphabet", "a-25-byte-long-string-val", "mid"]
    descs = [(1, 4), (0, 25), (2, 4), (0, 5), (1, 4)]
    cols = [rng.integers(-50, 50, n).astype(np.int32),
            oracle.pack_strings([words[i] for i in rng.integers(0, len(words), n)], 25),
            (rng.integers(-400, 400, n) / 8).astype(np.float32),
            oracle.pack_strings([["x", "yy", "zzz", "abcde", ""][i] for i in rng.integers(0, 5, n)], 5),
            rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32)]
    return descs, cols


@pytest.mark.parametrize("keys,desc", [([0], False), ([0], True), ([1], False), ([1], True), ([3, 0], False), ([2], False), ([2], True),
                                       ([4], False), ([1, 3, 0, 2], True), ([0, 4], False)])
def test_sort_against_oracle(ctx, oracle, keys, desc):
    n = 30_011
    descs, cols = _table(oracle, n, 3)
    t = load_table(ctx, descs, cols)
    exp = oracle.sort(descs, cols, keys, descending=desc)
    res = t.sort(keys, descending=desc, proj=[1, 0, 3], want=ALL)
    assert res.count == n
    np.testing.assert_array_equal(res.positions(), exp)
    np.testing.assert_array_equal(res.column(1), cols[0][exp])
    np.testing.assert_array_equal(res.column(0), cols[1][exp])
    np.testing.assert_array_equal(res.column(2), cols[3][exp])
    scan = oracle.scan(descs, cols, [], proj=[1, 0, 3])                      # clean Tuple bytes of every row, by position
    np.testing.assert_array_equal(res.tuples(), scan["tuples"][exp])
    res.close(); t.close()


def test_sort_skips_deleted_rows_and_handles_small_tables(ctx, oracle):
    descs, cols = _table(oracle, 5000, 5)
    t = load_table(ctx, descs, cols)
    dele = [0, 7, 4999, 1234]
    t.set_deleted(oracle.bits_from_positions(np.array(dele), 5000))
    res = t.sort([3, 1], proj=[3], want=ALL)
    np.testing.assert_array_equal(res.positions(), oracle.sort(descs, cols, [3, 1], deleted_positions=dele))
    assert res.count == 4996
    res.close(); t.close()
    for n in (0, 1, 2):
        d2, c2 = _table(oracle, n, 1)
        t2 = load_table(ctx, d2, c2)
        r2 = t2.sort([1, 0], descending=True, proj=[0], want=ALL)
        assert r2.count == n
        np.testing.assert_array_equal(r2.positions(), oracle.sort(d2, c2, [1, 0], descending=True))
        r2.close(); t2.close()
    with pytest.raises(mbcol.MbcError):
        load_table(ctx, descs, cols).sort([9])


def test_sort_golden_through_the_mirror(ctx, oracle, minidata, golden, tmp_path):
    """input.ColumnarSort.execute on the reference's own table: the six `sort` runs of the transcript, printed exactly as
    the Java printed them (the mirror replays the external merge sort's page / run structure over the key groups of the
    GPU result to order equal keys)."""
    from mbcol.global_ import SystemDefs
    from mbcol.input import ColumnarSort
    names, descs, cols = minidata
    w = oracle.DBWriter()
    oracle.write_columnar_file(w, "cf", names, descs, cols)
    path = str(tmp_path / "db")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    SystemDefs(path, 0, 100, None)
    try:
        n = 0
        for e in golden:
            if e["kind"] != "sort" or e.get("failed"):
                continue
            parts = e["cmd"].split()
            q = ColumnarSort()
            lines = q.execute(parts[1:], echo=False)
            assert lines[0] == "SORTED COLUMNS" and lines[-1] == "500" and q.resultCount == 500
            assert lines[1:-1] == e["rows"], e["cmd"]              # line by line, order of equal keys included
            n += 1
        assert n == 6
        assert ColumnarSort().execute(["db", "cf", "[A]", "[A]", "ASC", "16", "2"], echo=False)[0].startswith("NUMBUF_SORT is less than 3")
    finally:
        SystemDefs.shutdown()
