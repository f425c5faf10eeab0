"""K1 (mbc_table_ingest_dbfile) on images the Java could have left behind after deletes and a purge: interior EMPTY slots,
an emptied directory slot, markedDeleted bits -- none of those positions may come back as a live (zero-valued) row
(columnar/Columnarfile.java:874,912-914 purgeAllDeletedTuples deletes the heap records, then clears the bits; heap/Scan and
columnar/TupleScan never return an empty slot).  Needs a B200."""
import struct

import numpy as np
import pytest

import mbcol
from mbcol import _native as N

pytestmark = pytest.mark.gpu

PAGE, DP, PER_DIR = 1024, 20, 83


def _data_page_of(oracle, img, name, col, page_index):
    """(directory page id, directory slot, data page id) of the page_index-th data page of <name>.<col>."""
    dpid = oracle._file_entries(img)[f"{name}.{col}"]
    for _ in range(page_index // PER_DIR):
        dpid = struct.unpack_from(">i", img, dpid * PAGE + 12)[0]
    slot = page_index % PER_DIR
    ln, off = struct.unpack_from(">hH", img, dpid * PAGE + DP + 4 * slot)
    assert ln == 8
    return dpid, slot, struct.unpack_from(">i", img, dpid * PAGE + off + 4)[0]


def test_purged_rows_do_not_come_back(ctx, oracle):
    n = 30_011
    descs = [(1, 4), (0, 16), (2, 4)]
    cols = [oracle.synth_int(3, 0, n, 1000) + 1, oracle.synth_str(3, 1, n, 16), oracle.synth_real(3, 2, n) + 1.0]   # no zero values
    marked = [4, 999, 20_000]
    w = oracle.DBWriter()
    oracle.write_columnar_file(w, "t", ["I", "S", "R"], descs, cols, deleted_positions=marked)
    img = bytearray(w.tobytes())
    # (a) interior empty slots: positions purged in EVERY column (what purgeAllDeletedTuples leaves)
    purged = [0, 1, 77, 124, 125, 5000, 5001, 29_999, n - 1]
    for c, (t, wd) in enumerate(descs):
        per_page = (PAGE - DP) // (4 + (wd + 2 if t == 0 else wd))
        for p in purged:
            _, _, data = _data_page_of(oracle, bytes(img), "t", c, p // per_page)
            struct.pack_into(">h", img, data * PAGE + DP + 4 * (p % per_page), -1)            # EMPTY_SLOT (HFPage.java:300)
    # (b) an emptied directory slot in the int column: its whole data page (125 positions) is gone
    gone_page = 17
    dpid, slot, _ = _data_page_of(oracle, bytes(img), "t", 0, gone_page)
    struct.pack_into(">h", img, dpid * PAGE + DP + 4 * slot, -1)
    gone = list(range(125 * gone_page, 125 * gone_page + 125))
    dead = sorted(set(marked) | set(purged) | set(gone))
    live = np.setdiff1d(np.arange(n), dead)

    t = ctx.ingest_dbfile(bytes(img), "t")
    assert t.nrows == n - 1                                         # the row count is bounded by the highest OCCUPIED slot: the last row was purged
    res = t.scan([], proj=[0, 1, 2], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_HOST, aggs=[(0, 0), (1, 0)])
    np.testing.assert_array_equal(res.positions(), live)
    np.testing.assert_array_equal(res.column(0), cols[0][live])
    np.testing.assert_array_equal(res.column(1), cols[1][live])
    np.testing.assert_array_equal(res.column(2).view(np.uint32), cols[2][live].view(np.uint32))
    assert res.agg(0)[0] == live.size and res.agg(1)[0] == int(cols[0][live].astype(np.int64).sum())
    res.close()
    # a predicate that the zero fill of a purged position would satisfy selects nothing there
    z = t.scan([mbcol.Term(N.OP_EQ, ("col", 0), ("int", 0), 0)], want=N.WANT_POSITIONS | N.WANT_HOST)
    assert z.count == 0
    z.close()
    # the bitmap index skips them too (ColumnScan skips deleted rows at build)
    t.bitmap_build(0)
    assert 0 not in t.bitmap_values(0).tolist()
    t.close()


def test_image_without_holes_has_no_deleted_rows(ctx, oracle):
    n = 12_345
    descs = [(1, 4), (0, 7)]
    cols = [oracle.synth_int(5, 0, n, 50), oracle.pack_strings(["v%d" % (i % 97) for i in range(n)], 7)]
    w = oracle.DBWriter()
    oracle.write_columnar_file(w, "u", ["A", "B"], descs, cols)
    t = ctx.ingest_dbfile(w.tobytes(), "u")
    res = t.scan([], proj=[0, 1], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_HOST)
    assert res.count == n
    np.testing.assert_array_equal(res.column(0), cols[0])
    np.testing.assert_array_equal(res.column(1), cols[1])
    res.close(); t.close()
