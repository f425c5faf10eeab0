"""The Java operator surface, mirrored (same class / method names as minijava/src), driven the way the
reference's own drivers and ad-hoc mains drive it (input/Query.java, input/BitMapQuery.java,
input/MultiIndexQuery.java, tests2/IndexTest.java), and checked against the golden transcript and the oracle.
Needs a B200."""
import os

import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from mbcol.columnar import Columnarfile
from mbcol.global_ import AttrOperator, AttrType, IndexType, IntegerValue, StringValue, SystemDefs, TID
from mbcol.index import ColumnarIndexScan, ColumnIndexScan
from mbcol.input import BatchInsert, BitMapQuery, DeleteQuery, Index, MultiIndexQuery, NljQuery, Query, build_cnf_condexpr
from mbcol.iterator import ColumnarColumnScan, ColumnarColumnsScan, ColumnarFileScan, CondExpr, FldSpec, RelSpec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def db(tmp_path_factory, oracle, minidata):
    """A reference-format DB file holding cf, cf1, cf2 (= minidata, like phase3_output:24355-24373) and a
    synthetic int/real/char table with deleted rows; opened through SystemDefs like every reference driver."""
    names, descs, cols = minidata
    w = oracle.DBWriter()
    for cf in ("cf", "cf1", "cf2"):
        oracle.write_columnar_file(w, cf, names, descs, cols)
    n = 20_011
    sd, sc = [(1, 4), (2, 4), (0, 16), (0, 5)], None
    sc = [oracle.synth_int(7, 0, n, 50), oracle.synth_real(7, 1, n), oracle.synth_str(7, 2, n, 16),
          oracle.pack_strings([["x", "yy", "zzz", "abcde"][i % 4] for i in range(n)], 5)]
    oracle.write_columnar_file(w, "syn", ["I", "R", "S", "T"], sd, sc, deleted_positions=[0, 5, 77, 12345, n - 1])
    path = str(tmp_path_factory.mktemp("db") / "db")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    SystemDefs.shutdown()
    SystemDefs(path, 0, 100, None)
    yield {"path": path, "syn": (sd, sc, n)}
    SystemDefs.shutdown()


def test_ingest_matches_the_page_decoder(db, oracle, minidata):
    """K1: GPU decode of the heap pages == the oracle's reader == the arrays that were written."""
    names, descs, cols = minidata
    cf = Columnarfile("cf")
    assert cf.getFieldCount() == 4 and cf.getAttrNames() == names and cf.getTupleCnt() == 500
    assert [t.attrType for t in cf.getAttributeTypes()] == [0, 0, 1, 1] and cf.getStringSizes() == [25, 25]
    img = open(db["path"], "rb").read()
    ref = oracle.read_columnar_file(img, "cf")
    for c in range(4):
        got = cf.table.read_column(c)
        np.testing.assert_array_equal(got, ref["columns"][c])
        np.testing.assert_array_equal(got, cols[c])
    sd, sc, n = db["syn"]
    syn = Columnarfile("syn")
    assert syn.getTupleCnt() == n
    for c in range(4):
        np.testing.assert_array_equal(syn.table.read_column(c).view(np.uint8), np.asarray(sc[c]).view(np.uint8))
    # markedDeleted came along: TupleScan skips those rows (TupleScan.java:85)
    assert syn.getMarkedDeleted().getBitSet().positions().tolist() == [0, 5, 77, 12345, n - 1]
    ts = syn.openTupleScan()
    tid, seen = TID(4), []
    while ts.getNext(tid) is not None:
        seen.append(tid.position)
    assert len(seen) == n - 5 and seen[:3] == [1, 2, 3] and 12345 not in seen
    with pytest.raises(Exception, match="Columnar File does not exist"):
        Columnarfile("nope")


def test_query_filescan_single_column_predicates(db, oracle, minidata):
    """BASELINE config C1: `query db cf [targets] {col,op,val} NUMBUF FILESCAN` for every operator."""
    names, descs, cols = minidata
    n = 0
    for col, lits in (("A", ["Colorado", "South_Dakota"]), ("C", ["6", "0"]), ("D", ["3"])):
        for op in ("=", "<", ">", "!=", "<=", ">="):
            for lit in lits:
                for targets in ("[A,B,C,D]", "[C,A]", "[D]"):
                    lines = Query().execute(["db", "cf", targets, "{%s,%s,%s}" % (col, op, lit), "100", "FILESCAN"], echo=False)
                    tcols = [names.index(x) for x in targets[1:-1].split(",")]
                    ci = names.index(col)
                    term = oracle.Term(oracle.OPS[op], ("col", ci), ("int", int(lit)) if descs[ci][0] == 1 else ("str", lit), 0)
                    exp = oracle.scan(descs, cols, [term], proj=tcols)
                    rows = [", ".join(str(v) for v in oracle.decode_tuple(bytes(t), [descs[c] for c in tcols])) for t in exp["tuples"]]
                    assert lines[0] == ", ".join(targets[1:-1].split(","))
                    assert lines[1:1 + len(rows)] == rows
                    assert f"Total Results Count By Query: {exp['count']}" in lines
                    n += 1
    assert n == 90


def test_columnarfilescan_iterator_contract(db, oracle, minidata):
    """Constructor arguments, get_next / get_next_tid / restart / close / getTupleSize, and the reused Jtuple:
    Convert.setStrValue writes len+2 bytes only, so the padding of a string slot keeps older bytes -- the mirror
    reproduces the reference's Jtuple byte for byte (oracle stale_padding mode)."""
    names, descs, cols = minidata
    cf = Columnarfile("cf")
    types, sizes = cf.getAttributeTypes(), cf.getStringSizes()
    expr = CondExpr()
    expr.op = AttrOperator(AttrOperator.aopGE)
    expr.type1, expr.type2 = AttrType(AttrType.attrSymbol), AttrType(AttrType.attrInteger)
    expr.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), 3)          # 1-based: column C
    expr.operand2.integer = 5
    proj = [FldSpec(RelSpec(RelSpec.outer), i) for i in (1, 2, 3, 4)]
    fs = ColumnarFileScan("cf", types, sizes, 4, 4, proj, [expr, None])
    exp = oracle.scan(descs, cols, [oracle.Term(oracle.OP_GE, ("col", 2), ("int", 5), 0)], proj=[0, 1, 2, 3], stale_padding=True)
    assert fs.getTupleSize() == exp["tuple_len"] == 74 and fs.show() is not None
    got = []
    while True:
        t = fs.get_next()
        if t is None:
            break
        got.append(t.getTupleByteArray())
        assert t is fs.Jtuple                                            # one reused tuple object, like the Java
    assert len(got) == exp["count"] > 0
    assert got == [bytes(x) for x in exp["tuples"]]
    assert fs.get_next() is None
    fs.restart()
    tids = []
    while True:
        tid = fs.get_next_tid()
        if tid is None:
            break
        tids.append(tid.position)
    assert tids == exp["positions"].tolist()
    fs.close(); fs.close()                                               # idempotent (closeFlag)
    # tid-only constructor (delete query) and p == null
    fs2 = ColumnarFileScan("cf", types, sizes, 4, None)
    assert sum(1 for _ in iter(fs2.get_next_tid, None)) == 500
    assert fs2.aggregate([(0, 0), (1, 2), (2, 3), (3, 3)]) == [(500, True), (int(cols[2].sum()), True), (int(cols[3].min()), True), (int(cols[3].max()), True)]
    fs2.close()


def _cond(op, fld, lit):
    e = CondExpr()
    e.op = AttrOperator(op)
    e.type1 = AttrType(AttrType.attrSymbol)
    e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), fld)
    if isinstance(lit, str):
        e.type2, e.operand2.string = AttrType(AttrType.attrString), lit
    else:
        e.type2, e.operand2.integer = AttrType(AttrType.attrInteger), lit
    return e


def test_columnar_column_scans(db, oracle, minidata):
    """iterator.ColumnarColumnScan / ColumnarColumnsScan.  The CondExpr addresses the fields of the
    tuple of SCANNED columns (field k = colNos[k-1]), the output fields are fetched by position from out_indexes, the
    reused Jtuple is rewritten whole (Tuple.setFld of the stored record: clean padding, unlike ColumnarFileScan), deleted
    rows are skipped, and `query ... COLUMNSCAN` prints what FILESCAN prints."""
    names, descs, cols = minidata
    cf = Columnarfile("cf")
    # one scanned column: C >= 5 (field 1 of the predicate tuple), output [D, A]
    proj = [FldSpec(RelSpec(RelSpec.outer), 4), FldSpec(RelSpec(RelSpec.outer), 1)]
    cs = ColumnarColumnScan(cf, 2, 2, [3, 0], proj, [_cond(AttrOperator.aopGE, 1, 5), None])
    exp = oracle.scan(descs, cols, [oracle.Term(oracle.OP_GE, ("col", 2), ("int", 5), 0)], proj=[3, 0])
    got = []
    while (t := cs.get_next()) is not None:
        assert t is cs.Jtuple
        got.append(t.getTupleByteArray())
    assert len(got) == exp["count"] > 0 and got == [bytes(x) for x in exp["tuples"]]       # clean Tuple bytes
    cs.close(); cs.close()
    # two scanned columns [A, D]: (A <= "Delaware" OR D = 3) AND A != "Alabama"; fields 1 and 2 of the predicate tuple
    e1 = _cond(AttrOperator.aopLE, 1, "Delaware")
    e1.next = _cond(AttrOperator.aopEQ, 2, 3)
    e2 = _cond(AttrOperator.aopNE, 1, "Alabama")
    proj = [FldSpec(RelSpec(RelSpec.outer), i) for i in (2, 3, 1)]
    ms = ColumnarColumnsScan(cf, [0, 3], 3, [1, 2, 0], proj, [e1, e2, None])
    terms = [oracle.Term(oracle.OP_LE, ("col", 0), ("str", "Delaware"), 0), oracle.Term(oracle.OP_EQ, ("col", 3), ("int", 3), 0),
             oracle.Term(oracle.OP_NE, ("col", 0), ("str", "Alabama"), 1)]
    exp = oracle.scan(descs, cols, terms, proj=[1, 2, 0])
    got = [t.getTupleByteArray() for t in iter(ms.get_next, None)]
    assert 0 < exp["count"] < 500 and got == [bytes(x) for x in exp["tuples"]]
    ms.close()
    # delete-query constructors: positions only; a field number beyond the scanned columns is an error
    dq = ColumnarColumnsScan(cf, [0, 3], [e1, e2, None])
    assert [tid.position for tid in iter(dq.get_next_tid, None)] == exp["positions"].tolist()
    dq.close()
    dq1 = ColumnarColumnScan(cf, 3, [_cond(AttrOperator.aopEQ, 1, 3), None])
    assert [tid.position for tid in iter(dq1.get_next_tid, None)] == np.flatnonzero(cols[3] == 3).tolist()
    dq1.close()
    with pytest.raises(Exception):
        ColumnarColumnScan(cf, 3, [_cond(AttrOperator.aopEQ, 2, 3), None]).get_next_tid()
    # deleted rows are skipped (ColumnScan.getNext), on the table with deleted positions {0, 5, 77, 12345, n-1}
    sd, sc, n = db["syn"]
    syn = Columnarfile("syn")
    ss = ColumnarColumnScan(syn, 0, [_cond(AttrOperator.aopGE, 1, 0), None])
    seen = [tid.position for tid in iter(ss.get_next_tid, None)]
    assert len(seen) == n - 5 and seen[:3] == [1, 2, 3] and 12345 not in seen
    ss.close()
    # the driver: COLUMNSCAN prints the same lines as FILESCAN
    for cons in ("{C,>=,6}", "{A,<,Delaware}", "{D,!=,3}"):
        a = Query().execute(["db", "cf", "[B,D,A]", cons, "100", "COLUMNSCAN"], echo=False)
        b = Query().execute(["db", "cf", "[B,D,A]", cons, "100", "FILESCAN"], echo=False)
        assert a == b and len(a) > 3


def test_index_and_bitmap_accessors(db, oracle, minidata, golden):
    names, descs, cols = minidata
    for cfname in ("cf", "cf1", "cf2"):
        for col in "ABCD":
            Index().createIndex(["db", cfname, col, "bitmap"])
    cf = Columnarfile("cf")
    assert all(cf.bitmapIndexExists(c) for c in range(4))
    assert cf.getBitmapValues(2) == set(range(10)) and "South_Dakota" in cf.getBitmapValues(0)
    for e in golden:                                                     # G13/G14
        if e["kind"] == "index" and e.get("bitmap_bytes") and e["cmd"].split()[2] == "cf":
            c = names.index(e["cmd"].split()[3])
            sizes = [len(cf.getBitmapIndex(c, v).getBitSet().toByteArray()) for v in sorted(cf.getBitmapValues(c))]
            assert sorted(sizes) == sorted(e["bitmap_bytes"])
    bs = cf.getBitmapIndex(3, IntegerValue(3)).getBitSet()
    assert bs.positions().tolist() == np.nonzero(cols[3] == 3)[0].tolist()
    assert bs.nextSetBit(0) == int(np.nonzero(cols[3] == 3)[0][0]) and bs.cardinality() == int((cols[3] == 3).sum())
    assert cf.getBitmapIndex(0, StringValue("Atlantis")).getBitSet().isEmpty()


def test_indexes_query_golden(db, golden):
    """G10-G12 through MultiIndexQuery -> ColumnarIndexScan: the exact rows, in order."""
    n = 0
    for e in golden:
        if e["kind"] != "indexes_query" or e.get("failed"):
            continue
        args = e["cmd"].split()[1:]
        q = MultiIndexQuery()
        lines = q.execute(args, echo=False)
        assert q.resultCount == e["count"], e["cmd"]
        assert lines[0] == e["header"] and lines[1:1 + e["count"]] == e["rows"], e["cmd"]
        n += 1
    assert n >= 5


def test_bmj_golden(db, golden):
    """G2-G4, G6, G7, G9 through BitMapQuery.execute: printed bitsets, rows, order and count."""
    seen, n = set(), 0
    for e in golden:
        if e["kind"] != "bmj" or e.get("failed") or e["cmd"] in seen:
            continue
        seen.add(e["cmd"])
        q = BitMapQuery()
        lines = q.execute(e["cmd"].split()[1:], echo=False)
        assert q.resultCount == e["count"], e["cmd"]
        assert lines[1] == "{" + ", ".join(map(str, e["outer_bitset"])) + "}"
        assert lines[3] == "{" + ", ".join(map(str, e["inner_bitset"])) + "}"
        assert lines[4] == e["header"] and lines[5:5 + e["count"]] == e["rows"], e["cmd"]
        n += 1
    assert n >= 8


def test_nlj_golden(db, golden):
    """G5, G8 through NljQuery.execute: every `nlj` command of the transcript on cf/cf1/cf2, any
    access path the GPU side serves (FILESCAN / COLUMNSCAN / BITMAP).  Count, header and the rows IN ORDER must match what
    the Java printed: the mirror puts the GPU pair list into ColumnarNestedLoopJoins' block order."""
    import hashlib
    seen, n = set(), 0
    for e in golden:
        if e["kind"] != "nlj" or e.get("failed") or e["cmd"] in seen or "ff1." in e["cmd"] or "BTREE" in e["cmd"]:
            continue
        seen.add(e["cmd"])
        q = NljQuery()
        lines = q.execute(e["cmd"].split()[1:], echo=False)
        assert q.resultCount == e["count"], e["cmd"]
        assert lines[0] == e["header"], e["cmd"]
        rows = lines[1:1 + e["count"]]
        if "rows" in e and isinstance(e["rows"], list):
            assert rows == e["rows"], e["cmd"]                                  # line by line, in the Java's order
        assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == e["rows_sha256"], e["cmd"]
        if "BITMAP" not in e["cmd"]:
            # the same through two scan iterators and iterator.ColumnarNestedLoopJoins, built like NljQuery.java does
            assert NljQuery().execute(e["cmd"].split()[1:], echo=False, via_iterators=True) == lines, e["cmd"]
        n += 1
    assert n >= 8


def test_columnarindexscan_duplicate_constraint_cache(db, oracle, minidata):
    """The reference caches the CONJUNCT's accumulating BitSet for a repeated term (ColumnarIndexScan.java:147-172);
    the mirror's CNF rewrite gives the same positions as the oracle's literal emulation."""
    names, descs, cols = minidata
    cf = Columnarfile("cf")
    indexes = {c: oracle.bitmap_build(descs[c], cols[c]) for c in range(4)}
    for q in ["{(A,=,South_Dakota,BM)|(B,=,South_Dakota,BM)}^{(A,=,South_Dakota,BM)|(C,=,6,BM)}",
              "{(C,=,6,BM)|(C,=,6,BM)}^{(D,<,5,BM)}",
              "{(C,<,3,BM)}^{(D,=,1,BM)|(C,<,3,BM)}^{(C,<,3,BM)|(A,=,Colorado,BM)}",
              "{(A,=,Colorado,BM)|(B,=,Colorado,BM)}^{(C,!=,6,BM)}"]:
        exprs, itypes, fnums, inames = build_cnf_condexpr(q, cf)
        scan = ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), 4, exprs)
        exp = oracle.bitmap_cnf(indexes, names, oracle.parse_cnf(q, names, descs), 500, emulate_duplicate_cache=True)
        assert scan.getOutputPositions().positions().tolist() == oracle.positions_from_bits(exp, 500).tolist(), q
        scan.close()


def test_columnindexscan_single_predicate(db, oracle, minidata):
    """tests2/IndexTest.java:64-69 style: ColumnIndexScan(Bitmap, cf, name, types, sizes, n, selects, fldNum)."""
    names, descs, cols = minidata
    cf = Columnarfile("cf")
    for col, op, lit in ((2, "<=", 3), (3, "!=", 7), (0, ">", "Nevada"), (1, "=", "Delaware")):
        e = CondExpr()
        e.op = AttrOperator.findOperator(op)
        e.type1 = AttrType(AttrType.attrSymbol)
        e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), 1)
        if isinstance(lit, int):
            e.type2, e.operand2.integer = AttrType(AttrType.attrInteger), lit
        else:
            e.type2, e.operand2.string = AttrType(AttrType.attrString), lit
        s = ColumnIndexScan(IndexType(IndexType.Bitmap), cf, "", cf.getAttributeTypes(), cf.getStringSizes(), 4, [e, None], col + 1)
        idx = oracle.bitmap_build(descs[col], cols[col])
        exp = oracle.positions_from_bits(oracle.bitmap_term_bits(idx, oracle.OPS[op], lit, 500), 500).tolist()
        assert s.getPositionsOfIndexScan().positions().tolist() == exp
        assert [t.position for t in iter(s.get_next_tid, None)] == exp
        s.close()


def test_mark_deleted_is_seen_by_every_scan(db, oracle, minidata):
    names, descs, cols = minidata
    cf = Columnarfile("cf2")
    victims = [int(p) for p in np.nonzero(cols[2] == 6)[0][:7]]
    for p in victims:
        cf.markTupleDeleted(p)
    lines = Query().execute(["db", "cf2", "[C]", "{C,=,6}", "10", "FILESCAN"], echo=False)
    assert f"Total Results Count By Query: {int((cols[2] == 6).sum()) - 7}" in lines
    lines = Query().execute(["db", "cf2", "[C]", "{C,=,6}", "10", "BITMAP"], echo=False)
    assert f"Total Results Count By Query: {int((cols[2] == 6).sum()) - 7}" in lines
    for p in victims:
        cf.getMarkedDeleted().clear(p)
    assert "Total Results Count By Query: %d" % int((cols[2] == 6).sum()) in Query().execute(["db", "cf2", "[C]", "{C,=,6}", "10", "FILESCAN"], echo=False)


def test_delete_query_marks_rows_for_every_later_scan(db, oracle, minidata):
    """input.DeleteQuery (`delete_query DB CF {constraint} NUMBUF access md`): the TIDs come from get_next_tid of the
    tid-only scan constructors, markTupleDeleted hides the rows from every scan that follows -- through each access path,
    on a private copy of the table so the other tests keep theirs."""
    names, descs, cols = minidata
    from mbcol.global_ import AttrType as AT
    cf = Columnarfile("delq", 4, [AT(t) for t, _ in descs], [w for _, w in descs], names)
    cf.load_columns(cols)
    Index().createIndex(["db", "delq", "C", "bitmap"])
    gone = np.zeros(500, dtype=bool)
    for cons, access in (("{C,=,6}", "FILESCAN"), ("{A,<,Colorado}", "COLUMNSCAN"), ("{C,=,0}", "BITMAP")):
        col, op, lit = cons[1:-1].split(",")
        ci = names.index(col)
        term = oracle.Term(oracle.OPS[op], ("col", ci), ("int", int(lit)) if descs[ci][0] == 1 else ("str", lit), 0)
        hit = np.zeros(500, dtype=bool)
        hit[oracle.scan(descs, cols, [term])["positions"]] = True
        q = DeleteQuery()
        lines = q.execute(["db", "delq", cons, "100", access, "md"], echo=False)
        assert q.deletedCount == int((hit & ~gone).sum()) > 0                # rows deleted earlier are not seen again
        gone |= hit
        assert lines[1] == "500"
        assert cf.getMarkedDeleted().getBitSet().positions().tolist() == np.flatnonzero(gone).tolist()
        rest = Query().execute(["db", "delq", "[A,C]", "{D,>=,0}", "100", "FILESCAN"], echo=False)
        assert f"Total Results Count By Query: {int((~gone).sum())}" in rest
    assert Query().execute(["db", "delq", "[C]", "{C,=,6}", "100", "COLUMNSCAN"], echo=False)[-3].endswith(": 0")
    with pytest.raises(Exception, match="stays in Java"):
        DeleteQuery().execute(["db", "delq", "{C,=,1}", "100", "FILESCAN", "pd"], echo=False)
    cf.close()


def test_batchinsert_then_query_is_config_c1(db, golden):
    """BASELINE config C1 end to end through the drivers: `batchinsert minidata.txt db cfb 4` (G1: 500 records), then the
    same single-column FILESCAN queries as on the table ingested from the reference-format DB image."""
    bi = BatchInsert()
    lines = bi.insert([os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "minidata.tsv"), "db", "cfb", "4"], echo=False)
    rc = [e["record_count"] for e in golden if e["kind"] == "batchinsert" and "record_count" in e]
    assert lines == ["Record count: 500"] and bi.recordCount == 500 and 500 in rc
    for cons in ("{C,>=,6}", "{A,=,Colorado}", "{D,<,3}"):
        for access in ("FILESCAN", "COLUMNSCAN"):
            a = Query().execute(["db", "cfb", "[A,B,C,D]", cons, "100", access], echo=False)
            b = Query().execute(["db", "cf", "[A,B,C,D]", cons, "100", access], echo=False)
            assert a == b and len(a) > 6
    Columnarfile("cfb").close()
