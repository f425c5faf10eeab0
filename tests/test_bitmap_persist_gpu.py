"""SURVEY 8f rank 2 end to end: bitmap index built on the GPU (K3) -> BMIndexPage chains + catalogue written into the DB image
by the product (dbfile.persist_bitmap_index through the Columnarfile.createBitMapIndex mirror) -> read back by the oracle's
restatement of the Java reader (BitMapFile(String) / BM.readBitSet) == the oracle's own index build; the flushed DB file
re-opens with the indexes catalogued.  Needs a B200."""
import numpy as np
import pytest

import mbcol
from mbcol.columnar import Columnarfile
from mbcol.global_ import SystemDefs

pytestmark = pytest.mark.gpu


def test_gpu_built_index_persists_in_the_reference_format(tmp_path, oracle, minidata, golden):
    names, descs, cols = minidata
    w = oracle.DBWriter()
    oracle.write_columnar_file(w, "cf", names, descs, cols)
    n = 20_011
    sdescs = [(1, 4), (0, 16), (0, 5)]
    scols = [oracle.synth_int(7, 0, n, 50), oracle.synth_str(7, 2, n, 16),
             oracle.pack_strings([["x", "yy", "zzz", "abcde"][i % 4] for i in range(n)], 5)]
    dele = [0, 5, 77, 12345, n - 1]
    oracle.write_columnar_file(w, "syn", ["I", "S", "T"], sdescs, scols, deleted_positions=dele)
    path = str(tmp_path / "db")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    SystemDefs.shutdown()
    sd = SystemDefs(path, 0, 100, None)
    try:
        cf, syn = Columnarfile("cf"), Columnarfile("syn")
        for c in range(4):
            assert cf.createBitMapIndex(c) and cf.bitmapIndexExists(c)
        assert syn.createBitMapIndex(0) and syn.createBitMapIndex(2)
        img = sd.db_bytes
        assert oracle.space_map_pages(img) == set(range(len(img) // 1024))
        # minidata: every column, against the oracle's build and the transcript's byte lengths (G13 / G14)
        for c in range(4):
            exp = oracle.bitmap_build(descs[c], cols[c])
            back = oracle.read_bitmap_index(img, "cf", c)
            assert sorted(map(str, back)) == sorted(map(str, exp))
            for v, words in exp.items():
                np.testing.assert_array_equal(oracle.positions_from_bits(back[v], 500), oracle.positions_from_bits(words, 500))
            for e in golden:
                if e["kind"] == "index" and e.get("bitmap_bytes") and e["cmd"].split()[2] == "cf" and e["cmd"].split()[3] == names[c]:
                    assert sorted(oracle.bitset_bytearray_len(b) for b in back.values()) == sorted(e["bitmap_bytes"])
        # 20 011 rows with deleted positions: multi-page chains; deleted rows carry no bit (ColumnScan skips them at build)
        dwords = oracle.bits_from_positions(dele, n)
        for c in (0, 2):
            exp = oracle.bitmap_build(sdescs[c], scols[c], dwords)
            back = oracle.read_bitmap_index(img, "syn", c)
            assert sorted(map(str, back)) == sorted(map(str, exp))
            for v, words in exp.items():
                np.testing.assert_array_equal(oracle.positions_from_bits(back[v], n), oracle.positions_from_bits(words, n))
        assert oracle.read_bitmap_catalogue(img, "syn")["bitmapExist"] == [1, 0, 1]
        # the columns and markedDeleted survive the edit
        again = oracle.read_columnar_file(img, "syn")
        np.testing.assert_array_equal(again["columns"][0], scols[0])
        np.testing.assert_array_equal(oracle.positions_from_bits(again["deleted"], n), sorted(dele))
        # flush + reopen: the catalogue says which indexes exist
        sd.flush()
        SystemDefs.shutdown()
        SystemDefs(path, 0, 100, None)
        syn2 = Columnarfile("syn")
        assert syn2.bitmapIndexExists(0) and syn2.bitmapIndexExists(2) and not syn2.bitmapIndexExists(1)
        assert syn2.getBitmapValues(2) == {"x", "yy", "zzz", "abcde"}
    finally:
        SystemDefs.shutdown()
