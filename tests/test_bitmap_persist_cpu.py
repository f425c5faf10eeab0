"""Bitmap persistence in the reference's page format, host side: the product's DB image editor
(`dbfile.DBImage`, `dbfile.persist_bitmap_index`) writes what columnar/Columnarfile.java:698-753 + bitmap/BM.java:64-129
write, and the ORACLE's restatement of the Java reader (BitMapFile(String) -> BM.readBitSet, Columnarfile(String)'s
catalogue loop) reads it back.  Pinned on the reference's own transcript: the page numbers `batchinsert` writes and the
"Size while writing:N Pages to be written:M" lines of `index ... bitmap`.  No GPU: the bitmaps come from the oracle here;
tests/test_bitmap_gpu.py feeds the same writer from the device-built index."""
import os
import re

import numpy as np
import pytest

import mbcol
from mbcol import dbfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _minidata_image(oracle, minidata, deleted=()):
    names, descs, cols = minidata
    db = oracle.DBWriter()
    oracle.write_columnar_file(db, "cf", names, descs, cols, deleted)
    return db.tobytes()


def _bitset_bytes(words):
    return np.asarray(words, dtype=np.uint64).view(np.uint8).tobytes()


def test_batchinsert_page_numbers_match_the_transcript(oracle, minidata, golden):
    """phase3_output:15-20: `batchinsert minidata.txt db cf 4` wrote pages {0, 1, 129..176}: page 0 (file directory), page 1
    (the space-map page holding bits 0..8191 of a 1 M-page DB: 128 map pages) and 48 file pages.  The page writer allocates
    exactly those, and marks exactly those (plus the 128 map pages) in the space map."""
    img = _minidata_image(oracle, minidata)
    first = 1 + 128
    assert len(img) // 1024 == first + 48                               # pages 129..176
    used = oracle.space_map_pages(img)
    assert used == set(range(first + 48))
    entry = next((e for e in golden if e["kind"] == "batchinsert" and e.get("wrote_pages")), None)
    if entry is not None:                                               # the golden file carries the transcript's page map
        wrote = {int(k) for k in entry["wrote_pages"]}
        assert wrote - {0, 1} == set(range(first, first + 48))


def test_space_map_first_fit_and_directory_growth(oracle, minidata):
    img = dbfile.DBImage(_minidata_image(oracle, minidata))
    top = len(img.b) // 1024
    assert img.allocate_page() == top and img.allocate_page() == top + 1 and img.page_is_allocated(top + 1)
    # page 0 holds (1024 - 20) / 56 = 17 entries; 7 are taken (cf.hdr, cf.0..3, cf.md, cf.dtid): the 11th new file chains a
    # DBDirectoryPage with (1024 - 16) / 56 = 18 entries (diskmgr/DB.java:380-495)
    for k in range(30):
        img.add_file_entry(f"f{k}", 500 + k)
    files = dbfile._file_entries(img.b)
    assert all(files[f"f{k}"] == 500 + k for k in range(30)) and files["cf.hdr"] == 129
    nxt = int.from_bytes(img.b[0:4], "big", signed=True)
    assert nxt == top + 2 and int.from_bytes(img.b[nxt * 1024:nxt * 1024 + 4], "big", signed=True) == top + 3   # 10 + 18 + 2
    with pytest.raises(Exception, match="already exists"):
        img.add_file_entry("f3", 1)
    with pytest.raises(Exception, match="too long"):
        img.add_file_entry("x" * 50, 1)


@pytest.mark.parametrize("col", [0, 1, 2, 3])
def test_persisted_index_reads_back_through_the_java_reader(oracle, minidata, golden, col):
    """index db cf <col> bitmap: every value's file, the header catalogue and bitmapExist; golden G13/G14: the per-value
    BitSet.toByteArray() lengths ("Size while writing:") and page counts fall out of the page image."""
    names, descs, cols = minidata
    img0 = _minidata_image(oracle, minidata)
    index = oracle.bitmap_build(descs[col], cols[col])
    values = sorted(index, key=lambda v: v.encode() if isinstance(v, str) else v)
    img1 = dbfile.persist_bitmap_index(img0, "cf", col, values, [_bitset_bytes(index[v]) for v in values])
    cat = oracle.read_bitmap_catalogue(img1, "cf")
    assert cat["bitmapExist"] == [1 if c == col else 0 for c in range(4)]
    assert sorted(map(str, cat["values"][col])) == sorted(map(str, values)) and all(not cat["values"][c] for c in range(4) if c != col)
    back = oracle.read_bitmap_index(img1, "cf", col)
    for v in values:
        n = max(back[v].size, index[v].size)
        a, b = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        a[:back[v].size], b[:index[v].size] = back[v], index[v]
        np.testing.assert_array_equal(a, b, err_msg=f"bitmap of {v!r}")
    # the columns themselves are untouched and the space map covers every page of the new image
    again = oracle.read_columnar_file(img1, "cf")
    for c in range(4):
        np.testing.assert_array_equal(again["columns"][c], np.asarray(cols[c]).reshape(again["columns"][c].shape))
    assert oracle.space_map_pages(img1) == set(range(len(img1) // 1024))
    # transcript: one "Size while writing:<bytes>Pages to be written:<pages>" line per value
    for e in golden:
        if e["kind"] == "index" and e.get("bitmap_bytes") and e["cmd"].split()[3] == names[col]:
            sizes = sorted(oracle.bitset_bytearray_len(back[v]) for v in values)
            assert sizes == sorted(e["bitmap_bytes"])
            assert all((s + 999) // 1000 == 1 for s in sizes)           # 500 rows: every bitmap fits one 1000-byte record
    # transcript, first `index db cf A bitmap` (run on the file as batchinsert left it; the sort before it dropped its temp
    # files): it wrote page 0 (file directory), page 1 (space map), 129 / 130 (cf.hdr directory + data page) and 21 NEW pages
    # for 20 values: one header page per value plus the DBDirectoryPage chained when page 0's 17 slots ran out
    if col == 0:
        first = next(e for e in golden if e["cmd"] == "index db cf A bitmap" and e.get("wrote_pages"))
        new_in_transcript = [int(k) for k in first["wrote_pages"] if int(k) not in (0, 1, 129, 130)]
        assert len(new_in_transcript) == len(img1) // 1024 - len(img0) // 1024 == len(values) + 1 == 21
        assert max(new_in_transcript) - min(new_in_transcript) + 1 == 21        # one first-fit run, like ours (177..197)
        assert {int(k) for k in first["wrote_pages"]} >= {0, 1, 129, 130}
        changed = {p for p in range(len(img0) // 1024) if img0[p * 1024:(p + 1) * 1024] != img1[p * 1024:(p + 1) * 1024]}
        assert changed == {0, 1, 129, 130}                              # the same four existing pages the Java dirtied
    # a second call is a no-op (`if (bitmapExist[columnNo] != 1)`, Columnarfile.java:699)
    assert dbfile.persist_bitmap_index(img1, "cf", col, values, [_bitset_bytes(index[v]) for v in values]) == img1


def test_long_bitmaps_chain_pages_and_many_values_grow_the_header(oracle):
    """20 011 rows: a value's bitmap is 2 502 bytes = 3 BMIndexPages (1000 + 1000 + 502 zero padded); 300 distinct values
    overflow the header's data page (new data pages + DataPageInfo records) and the file directory (new directory pages)."""
    nrows = 20_011
    rng = np.random.default_rng(1)
    k = rng.integers(0, 300, nrows).astype(np.int32)
    g = rng.integers(0, 3, nrows).astype(np.int32)
    descs, names = [(1, 4), (1, 4)], ["K", "G"]
    db = oracle.DBWriter()
    oracle.write_columnar_file(db, "t", names, descs, [k, g], deleted_positions=[5, 77, 20_000])
    img = db.tobytes()
    for col, column in ((0, k), (1, g)):
        index = oracle.bitmap_build(descs[col], column)
        values = sorted(index)
        img = dbfile.persist_bitmap_index(img, "t", col, values, [_bitset_bytes(index[v]) for v in values])
        back = oracle.read_bitmap_index(img, "t", col)
        assert sorted(back) == values
        for v in values:
            np.testing.assert_array_equal(oracle.positions_from_bits(back[v], nrows), np.nonzero(column == v)[0])
    files = dbfile._file_entries(img)
    chain, pid = [], files["t.bm.1.0"]
    while pid != -1:
        chain.append(pid)
        assert int.from_bytes(img[pid * 1024 + 20:pid * 1024 + 22], "big") == 1000      # one 1000-byte record per page
        pid = int.from_bytes(img[pid * 1024 + 12:pid * 1024 + 16], "big", signed=True)
    assert len(chain) == 3 and int.from_bytes(img[chain[0] * 1024 + 6:chain[0] * 1024 + 8], "big") == 13     # BMHEAD
    assert int.from_bytes(img[chain[1] * 1024 + 8:chain[1] * 1024 + 12], "big", signed=True) == chain[0]     # prev pointer
    cat = oracle.read_bitmap_catalogue(img, "t")
    assert cat["bitmapExist"] == [1, 1] and len(cat["values"][0]) == 300 and sorted(cat["values"][1]) == [0, 1, 2]
    again = oracle.read_columnar_file(img, "t")
    np.testing.assert_array_equal(again["columns"][0], k)
    np.testing.assert_array_equal(oracle.positions_from_bits(again["deleted"], nrows), [5, 77, 20_000])
    assert oracle.space_map_pages(img) == set(range(len(img) // 1024))


def test_string_values(oracle, minidata):
    names, descs, cols = minidata
    img0 = _minidata_image(oracle, minidata)
    index = oracle.bitmap_build(descs[0], cols[0])
    values = sorted(index, key=lambda s: s.encode())
    img1 = dbfile.persist_bitmap_index(img0, "cf", 0, values, [_bitset_bytes(index[v]) for v in values])
    assert "cf.bm.0." + values[0] in dbfile._file_entries(img1)
    hdr = dbfile.read_header(img1, "cf")
    assert hdr["bitmapExist"][0] == 1 and sorted(hdr["bitmapValues"]) == sorted("0." + v for v in values)
