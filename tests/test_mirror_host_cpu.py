"""Host-side logic of the Java mirror that needs no GPU: heap.Tuple byte images against the oracle's Tuple encoder,
CondExpr -> term program, java.util.BitSet semantics, the DB-file header reader against the oracle's page writer."""
import numpy as np
import pytest

from mbcol.bitmap import BitSet
from mbcol.dbfile import read_header
from mbcol.global_ import AttrOperator, AttrType, TID
from mbcol.heap import FieldNumberOutOfBoundException, Tuple
from mbcol.iterator import CondExpr, FldSpec, RelSpec, flatten_condexpr, setup_op_tuple


def _types(descs):
    return [AttrType(t) for t, _ in descs], [w for t, w in descs if t == AttrType.attrString]


def test_tuple_bytes_match_the_oracle_encoder(oracle):
    """Tuple.setHdr / set*Fld (heap/Tuple.java:369-440, global/Convert.java) produce the byte image the oracle's row
    encoder produces for the same values, field offsets included."""
    descs = [(0, 25), (1, 4), (2, 4), (0, 5)]
    cols = [oracle.pack_strings(["Colorado", "x", ""], 25), np.array([7, -3, 2**31 - 1], np.int32),
            np.array([1.5, -0.25, 3e9], np.float32), oracle.pack_strings(["abcde", "", "yy"], 5)]
    exp = oracle.scan(descs, cols, [], proj=[0, 1, 2, 3])
    types, sizes = _types(descs)
    for row in range(3):
        t = Tuple()
        t.setHdr(4, types, sizes)
        assert t.size() == exp["tuple_len"] == len(exp["tuples"][row]) and t.noOfFlds() == 4
        t.setStrFld(1, oracle.unpack_strings(cols[0])[row]).setIntFld(2, int(cols[1][row])).setFloFld(3, float(cols[2][row]))
        t.setStrFld(4, oracle.unpack_strings(cols[3])[row])
        assert t.getTupleByteArray() == bytes(exp["tuples"][row])
        assert (t.getStrFld(1), t.getIntFld(2), t.getStrFld(4)) == (oracle.unpack_strings(cols[0])[row], int(cols[1][row]), oracle.unpack_strings(cols[3])[row])
        assert np.float32(t.getFloFld(3)) == cols[2][row]
        back = Tuple(t.getTupleByteArray())
        back._adopt_header()
        assert back.getIntFld(2) == int(cols[1][row]) and back.getStrFld(1) == t.getStrFld(1)
    with pytest.raises(FieldNumberOutOfBoundException):
        t.getIntFld(5)


def test_reused_tuple_keeps_stale_string_padding():
    """Convert.setStrValue writes [len][bytes] only: a shorter string leaves the tail of the longer one in the slot
    (what the reference's reused Jtuple does); setFld rewrites the whole slot."""
    t = Tuple()
    t.setHdr(1, [AttrType(AttrType.attrString)], [8])
    t.setStrFld(1, "alphabet")
    first = t.getTupleByteArray()
    t.setStrFld(1, "a")
    assert t.getStrFld(1) == "a"
    assert t.getTupleByteArray()[-8:] == b"a" + first[-7:]                     # "lphabet" is still there
    t.setFld(1, b"\x00\x01a" + b"\0" * 7)
    assert t.getTupleByteArray()[-8:] == b"a" + b"\0" * 7


def test_projection_header_and_condexpr_flattening():
    descs = [(0, 25), (0, 25), (1, 4), (1, 4)]
    types, sizes = _types(descs)
    j, out_types = Tuple(), [None, None]
    got = setup_op_tuple(j, out_types, types, 4, sizes, [FldSpec(RelSpec(RelSpec.outer), 3), FldSpec(RelSpec(RelSpec.outer), 2)], 2)
    assert [t.attrType for t in out_types] == [AttrType.attrInteger, AttrType.attrString] and got == [25]
    assert j.size() == (2 + 2) * 2 + 4 + 27

    def cond(op, fld, lit):
        e = CondExpr()
        e.op = AttrOperator(op)
        e.type1 = AttrType(AttrType.attrSymbol)
        e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), fld)
        if isinstance(lit, str):
            e.type2, e.operand2.string = AttrType(AttrType.attrString), lit
        elif isinstance(lit, float):
            e.type2, e.operand2.real = AttrType(AttrType.attrReal), lit
        else:
            e.type2, e.operand2.integer = AttrType(AttrType.attrInteger), lit
        return e

    a = cond(AttrOperator.aopLE, 1, "Delaware")
    a.next = cond(AttrOperator.aopEQ, 4, 3)                      # OR inside a conjunct: the .next chain
    b = cond(AttrOperator.aopGT, 3, 2.5)
    terms = flatten_condexpr([a, b, None, cond(AttrOperator.aopEQ, 1, "never reached")])
    assert [(t.op, t.lhs, t.rhs, t.conj) for t in terms] == [
        (AttrOperator.aopLE, ("col", 0), ("str", "Delaware"), 0), (AttrOperator.aopEQ, ("col", 3), ("int", 3), 0),
        (AttrOperator.aopGT, ("col", 2), ("real", 2.5), 1)]
    assert flatten_condexpr(None) == [] and flatten_condexpr([None]) == []
    assert AttrOperator.findOperator("<=").attrOperator == AttrOperator.aopLE
    assert AttrOperator.getOppositeOperator("<").attrOperator == AttrOperator.aopGT
    assert AttrOperator.getOppositeOperator("!=").attrOperator == AttrOperator.aopNE
    assert TID(4, 17).position == 17


def test_bitset_follows_java_util_bitset():
    b = BitSet()
    assert b.isEmpty() and b.length() == 0 and b.toByteArray() == b"" and repr(b) == "{}" and b.nextSetBit(0) == -1
    for p in (1, 2, 5, 64, 200):
        b.set(p)
    assert repr(b) == "{1, 2, 5, 64, 200}" and b.cardinality() == 5 and b.length() == 201
    assert len(b.toByteArray()) == 200 // 8 + 1                  # BitSet.toByteArray(): up to the highest set bit
    assert b.get(64) and not b.get(63) and not b.get(10_000)
    assert [b.nextSetBit(0), b.nextSetBit(3), b.nextSetBit(65), b.nextSetBit(201)] == [1, 5, 200, -1]
    b.clear(200)
    assert b.length() == 65 and b.positions().tolist() == [1, 2, 5, 64] and list(b) == [1, 2, 5, 64]
    c = BitSet()
    c.set(2); c.set(64); c.set(70)
    d = BitSet(b.toLongArray().copy())
    d.and_(c)
    assert d.positions().tolist() == [2, 64]
    d.or_(c)
    assert d.positions().tolist() == [2, 64, 70]
    d.andNot(b)
    assert d.positions().tolist() == [70] and d != c and BitSet(c.toLongArray().copy()) == c


def test_dbfile_header_reader_against_the_page_writer(oracle, minidata):
    names, descs, cols = minidata
    w = oracle.DBWriter()
    oracle.write_columnar_file(w, "cf", names, descs, cols, deleted_positions=[3, 64, 499])
    oracle.write_columnar_file(w, "other", ["K"], [(1, 4)], [np.arange(10, dtype=np.int32)])
    img = w.tobytes()
    h = read_header(img, "cf")
    assert h["numColumns"] == 4 and h["colnames"] == names
    assert BitSet(np.frombuffer(h["deleted_bytes"], dtype=np.uint64).copy()).positions().tolist() == [3, 64, 499]
    assert read_header(img, "other")["colnames"] == ["K"] and read_header(img, "other")["deleted_bytes"] == b""
    with pytest.raises(Exception, match="Columnar File does not exist"):
        read_header(img, "nope")


def test_batchinsert_datafile_parser(oracle, minidata):
    """input.parse_datafile (BatchInsert.java:60-103) against the fixture the oracle reads: names, types, column arrays."""
    import os
    from mbcol.input import parse_datafile
    names, descs, cols = minidata
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "minidata.tsv")
    n2, d2, c2 = parse_datafile(path, 4)
    assert n2 == names and [(int(t), w) for t, w in d2] == descs
    for a, b in zip(c2, cols):
        np.testing.assert_array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1))
    for bad in (0, 2, 5):                                        # NUMCOLUMNS must be the file's column count
        with pytest.raises(Exception, match="does not match"):
            parse_datafile(path, bad)


def test_host_orderings_match_the_oracle_restatements(oracle):
    """The two orderings the mirror applies on the host to GPU results -- the block nested-loop emission order of `nlj` and
    the reference's external-merge order of equal keys in `sort` -- against the oracle's restatements (which are pinned on
    the transcript), on random inputs."""
    from mbcol.input import _external_sort_replay, nlj_emission_order
    rng = np.random.default_rng(4)
    for _ in range(20):
        n_o, n_i = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        oq = np.sort(rng.choice(1000, n_o, replace=False))
        npairs = int(rng.integers(0, 2000))
        po, pi = oq[rng.integers(0, n_o, npairs)], rng.integers(0, n_i, npairs)
        block = int(rng.integers(1, 400))
        np.testing.assert_array_equal(nlj_emission_order(oq, po, pi, block), oracle.nlj_order(oq, po, pi, block))
    for trial in range(12):
        n = int(rng.integers(1, 4000))
        descs = [(1, 4), (0, 7)]
        cols = [rng.integers(0, int(rng.integers(1, 40)), n).astype(np.int32),
                oracle.pack_strings([["a", "bb", "ccc", ""][i] for i in rng.integers(0, 4, n)], 7)]
        keys = [[0], [1], [1, 0]][trial % 3]
        desc, bufs = bool(trial % 2), int(rng.integers(3, 9))
        exp = oracle.external_sort_order(descs, cols, keys, desc, bufs)
        stable = oracle.sort(descs, cols, keys, descending=desc)                  # what the GPU returns: ties by position
        change = np.zeros(max(n - 1, 0), dtype=bool)
        for c in keys:
            a = np.asarray(cols[c]).reshape(n, -1)[stable]
            change |= (a[1:] != a[:-1]).any(axis=1)
        group = np.concatenate([[0], np.cumsum(change)]) if n else np.zeros(0, dtype=np.int64)
        by_pos = np.argsort(stable, kind="stable")
        rec = sum(w + 2 if t == 0 else 4 for t, w in (descs[c] for c in keys)) + 4
        got = [int(stable[by_pos[r]]) for r in _external_sort_replay(group[by_pos].tolist(), 1004 // (rec + 4), bufs - 1)]
        assert got == [int(x) for x in exp], (trial, n, keys, desc, bufs)
