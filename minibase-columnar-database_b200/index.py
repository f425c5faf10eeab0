"""Mirror of the bitmap branches of index.ColumnIndexScan / index.ColumnarIndexScan
(minijava/src/index/ColumnIndexScan.java:185-272,600-740, ColumnarIndexScan.java:79-308).  The B-tree branches
stay in Java (serial pointer chasing; out of scope)."""
from __future__ import annotations

import struct
from typing import Optional, Sequence

from . import _native as N
from .bitmap import BitSet
from .engine import Term
from .global_ import AttrOperator, AttrType, IndexType, TID
from .heap import Tuple
from .iterator import (CondExpr, FldSpec, IndexException, Iterator, flatten_condexpr, setup_op_tuple)


def _slot_bytes(attr: int, width: int, value) -> bytes:
    """The raw column record a heapfile holds (what Jtuple.setFld copies, ColumnarIndexScan.java:287-308)."""
    if attr == AttrType.attrInteger:
        return struct.pack(">i", int(value))
    if attr == AttrType.attrReal:
        return struct.pack(">f", float(value))
    raw = bytes(value).rstrip(b"\0")
    return (struct.pack(">H", len(raw)) + raw).ljust(width + 2, b"\0")


class _GatherCursor:
    def __init__(self, cf, result, out_indexes, jtuple):
        self.cf, self.result, self.out_indexes, self.jtuple = cf, result, list(out_indexes), jtuple
        self.positions = result.positions()
        self.cols = [result.column(i) for i in range(len(self.out_indexes))]
        self.i = 0

    def next_tuple(self) -> Optional[Tuple]:
        if self.i >= len(self.positions):
            return None
        k = self.i
        self.i += 1
        for j, c in enumerate(self.out_indexes):
            self.jtuple.setFld(j + 1, _slot_bytes(self.cf.attrTypes[c].attrType, self.cf.attrSizes[c], self.cols[j][k]))
        return self.jtuple


class ColumnIndexScan(Iterator):
    """One predicate `column op literal` answered from the column's bitmap index.

    ColumnIndexScan(index, cf, indName, types, str_sizes, noInFlds, noOutFlds, out_indexes, outFlds, selects, fldNum, indexOnly)  (:76)
    ColumnIndexScan(index, cf, indName, types, str_sizes, noInFlds, selects, fldNum)                                            (:185)"""

    def __init__(self, index: IndexType, cf, indName, types, str_sizes, noInFlds, *rest):
        super().__init__()
        if len(rest) == 6:
            noOutFlds, out_indexes, outFlds, selects, fldNum, indexOnly = rest
        elif len(rest) == 2:
            selects, fldNum = rest
            noOutFlds, out_indexes, outFlds, indexOnly = 0, [], [], False
        else:
            raise TypeError("bad ColumnIndexScan arguments")
        if index.indexType != IndexType.Bitmap:
            raise IndexException(None, "Only bitmap index scans run on the GPU; B-tree scans stay in Java")
        self.f, self._noInFlds, self._selects, self.colNo = cf, noInFlds, selects, fldNum - 1
        self.outIndexes, self.index_only = list(out_indexes), indexOnly
        self.Jtuple = Tuple()
        if noOutFlds:
            setup_op_tuple(self.Jtuple, [None] * noOutFlds, types, noInFlds, str_sizes, outFlds, noOutFlds)
        sel = selects[0]                                        # only _selects[0] is used (:656-740)
        lit = ("int", sel.operand2.integer) if cf.attrTypes[self.colNo].attrType == AttrType.attrInteger else ("str", sel.operand2.string)
        try:
            self._result = cf.table.bitmap_scan([Term(sel.op.attrOperator, ("col", self.colNo), lit, 0)], proj=self.outIndexes,
                                                want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_BITMAP | N.WANT_HOST)
        except N.MbcError as e:
            raise IndexException(e, "IndexScan.java: " + e.message)
        self._cursor = _GatherCursor(cf, self._result, self.outIndexes, self.Jtuple)

    def get_next(self) -> Optional[Tuple]:
        return self._cursor.next_tuple()

    def get_next_tid(self) -> Optional[TID]:                     # get_bm_next_tid (:600-624): deleted rows skipped
        c = self._cursor
        if c.i >= len(c.positions):
            return None
        c.i += 1
        return TID(self._noInFlds, int(c.positions[c.i - 1]))

    def getPositionsOfIndexScan(self) -> BitSet:                 # :647-654
        return BitSet(self._result.bitmap())

    def close(self) -> None:
        if not self.closeFlag:
            self._result.close()
            self.closeFlag = True


class ColumnarIndexScan(Iterator):
    """CNF over per-term index scans, evaluated eagerly in the constructor like the reference (:130-181).

    ColumnarIndexScan(cf, fldNums, indexTypes, indNames, types, str_sizes, noInFlds, noOutFlds, out_indexes, outFlds, selects, indexOnly)  (:79)
    ColumnarIndexScan(cf, fldNums, indexTypes, indNames, types, str_sizes, noInFlds, selects)                                            (:185)

    The reference's duplicateConstraints cache (:147-172) is reproduced: a term whose text occurs more than once
    in the query is cached the first time it is evaluated, and what is cached is the conjunct's ACCUMULATING
    BitSet.  Because only ORs touch that object, the effect is a pure rewrite of the CNF, done here on the host
    before the single GPU call: a repeated term in a later conjunct stands for every term of the conjunct that
    evaluated it first."""

    def __init__(self, columnarFile, fldNums, indexTypes, indNames, types, str_sizes, noInFlds, *rest):
        super().__init__()
        if len(rest) == 5:
            noOutFlds, out_indexes, outFlds, selects, indexOnly = rest
        elif len(rest) == 1:
            (selects,) = rest
            noOutFlds, out_indexes, outFlds, indexOnly = 0, [], [], False
        else:
            raise TypeError("bad ColumnarIndexScan arguments")
        self.f, self._noInFlds, self._selects = columnarFile, noInFlds, selects
        self.outIndexes, self.index_only, self._noOutFlds = list(out_indexes), indexOnly, noOutFlds
        self.Jtuple = Tuple()
        if noOutFlds:
            setup_op_tuple(self.Jtuple, [None] * noOutFlds, types, noInFlds, str_sizes, outFlds, noOutFlds)
        terms = self._rewrite(selects)
        try:
            self._result = columnarFile.table.bitmap_scan(terms, proj=self.outIndexes,
                                                          want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_BITMAP | N.WANT_HOST)
        except N.MbcError as e:
            raise IndexException(e, "IndexScan.java: " + e.message)
        self._cursor = _GatherCursor(columnarFile, self._result, self.outIndexes, self.Jtuple)

    def _text(self, t: CondExpr) -> str:                         # :143-146
        lit = str(t.operand2.integer) if t.type2.attrType == AttrType.attrInteger else t.operand2.string
        return self.f.indexToColName(t.operand1.symbol.offset - 1) + t.op.toString() + lit + t.indexType.toString()

    def _rewrite(self, selects) -> list[Term]:
        conjuncts = []
        for head in selects:
            if head is None:
                break
            chain, cur = [], head
            while cur is not None:
                if not ((cur.type1.attrType == AttrType.attrSymbol) != (cur.type2.attrType == AttrType.attrSymbol)):
                    raise IndexException(None, "IndexScan.java: invalid constraint")          # :137-142
                chain.append(cur)
                cur = cur.next
            conjuncts.append(chain)
        query = "^".join("|".join(self._text(t) for t in conj) for conj in conjuncts)         # buildInputQueryString
        cache: dict[str, int] = {}                               # term text -> conjunct whose BitSet was cached
        final: list[list[CondExpr]] = []
        for i, conj in enumerate(conjuncts):
            mine: list[CondExpr] = []
            for t in conj:
                key = self._text(t)
                dup = key in cache or self._count(query, key) > 1
                if not dup or key not in cache:
                    mine.append(t)
                    if dup:
                        cache[key] = i
                elif cache[key] != i:                            # positions.or(cached): the whole earlier conjunct
                    mine.extend(final[cache[key]])
            final.append(mine)
        terms: list[Term] = []
        for i, conj in enumerate(final):
            for t in conj:
                col = t.operand1.symbol.offset - 1
                lit = ("int", t.operand2.integer) if self.f.attrTypes[col].attrType == AttrType.attrInteger else ("str", t.operand2.string)
                terms.append(Term(t.op.attrOperator, ("col", col), lit, i))
        return terms

    @staticmethod
    def _count(hay: str, needle: str) -> int:                    # checkDuplicateConstraint (:352-369)
        n, i = 0, 0
        while True:
            i = hay.find(needle, i)
            if i < 0:
                return n
            n += 1
            i += len(needle)

    def getOutputPositions(self) -> BitSet:
        return BitSet(self._result.bitmap())

    def get_next(self) -> Optional[Tuple]:                       # :287-308
        return self._cursor.next_tuple()

    def restart(self) -> None:
        self._cursor.i = 0

    def getTupleSize(self) -> int:
        return self.Jtuple.size()

    def close(self) -> None:
        if not self.closeFlag:
            self._result.close()
            self.closeFlag = True
