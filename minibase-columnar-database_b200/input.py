"""Mirror of the driver classes that sit directly on the scan path (minijava/src/input/): they marshal the
command-line arguments, build CondExpr[] / FldSpec[] and drive the operators.  Same argument order and the
same printed rows as the reference; every method also returns the rows so tests can assert on them."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .columnar import Columnarfile
from .engine import Term, bitmap_join
from .global_ import AttrOperator, AttrType, IndexType, SystemDefs
from .index import ColumnarIndexScan
from .iterator import ColumnarColumnScan, ColumnarFileScan, CondExpr, FldSpec, RelSpec


def _emit(lines: list, text: str, echo: bool) -> None:
    lines.append(text)
    if echo:
        print(text)


def _footer(lines: list, count: int, echo: bool) -> None:
    for t in ("", "*" * 72, f"Total Results Count By Query: {count}", "*" * 72, ""):
        _emit(lines, t, echo)


def build_cnf_condexpr(constraintStr: str, cf: Columnarfile, outer: bool = True):
    """BitMapQuery.buildCNFQueryCondExpr (:347-420) / MultiIndexQuery's 4-field form.
    `{(A,=,x)|(B,<,y,BM)}^{(C,!=,6)}` -> (CondExpr[] with trailing None, indexTypes, fldNums, indNames)."""
    exprs, index_types, fld_nums, ind_names = [], [], [], []
    for conj in constraintStr.split("^"):
        if not (conj.startswith("{") and conj.endswith("}")):
            raise Exception("Invalid query format")
        head = tail = None
        for dis in conj[1:-1].split("|"):
            if not (dis.startswith("(") and dis.endswith(")")):
                raise Exception("Invalid query format")
            parts = [p.strip() for p in dis[1:-1].strip().split(",")]
            if len(parts) not in (3, 4):
                raise Exception("Invalid VALUECONSTRAINT elements")
            col = cf.colNameToIndex(parts[0])
            t = CondExpr()
            t.op = AttrOperator.findOperator(parts[1])
            t.type1 = AttrType(AttrType.attrSymbol)
            t.operand1.symbol = FldSpec(RelSpec(RelSpec.outer if outer else RelSpec.innerRel), col + 1)
            if cf.getAttributeTypes()[col].attrType == AttrType.attrInteger:
                t.type2 = AttrType(AttrType.attrInteger)
                t.operand2.integer = int(parts[2])
            else:
                t.type2 = AttrType(AttrType.attrString)
                t.operand2.string = parts[2]
            access = parts[3] if len(parts) == 4 else "BM"
            t.indexType = IndexType(IndexType.Bitmap if access.upper() in ("BM", "BITMAP") else IndexType.B_Index)
            index_types.append(t.indexType)
            ind_names.append(" ")
            fld_nums.append(col + 1)
            if head is None:
                head = tail = t
            else:
                tail.next = t
                tail = t
        exprs.append(head)
    exprs.append(None)
    return exprs, index_types, fld_nums, ind_names


def _fmt(tuple_, types: Sequence[int]) -> str:
    vals = []
    for i, t in enumerate(types):
        vals.append(str(tuple_.getIntFld(i + 1)) if t == AttrType.attrInteger else
                    repr(tuple_.getFloFld(i + 1)) if t == AttrType.attrReal else tuple_.getStrFld(i + 1))
    return ", ".join(vals)


def parse_datafile(path: str, numcolumns: int):
    """The data-file format of input/BatchInsert.java:60-103: a header row `name:int` / `name:char(N)` separated by tabs,
    then one row per record.  Returns (names, [(attrType, size)], column arrays in the mbc_table_load_column layout)."""
    with open(path, "r", encoding="utf-8") as f:
        headers = f.readline().rstrip("\r\n").split("\t")
        if len(headers) != numcolumns or numcolumns == 0:        # (the Java indexes past the header array when NUMCOLUMNS is larger)
            raise Exception("Number of columns specified does not match the number of columns in data file")
        names, descs = [], []
        for h in headers[:numcolumns]:
            name, typ = h.split(":")
            names.append(name)
            if typ == "int":
                descs.append((AttrType.attrInteger, 4))
            elif typ.startswith("char"):
                descs.append((AttrType.attrString, int(typ[len("char("):-1])))
            else:
                raise Exception("column attr type is not supported.")
        rows = [ln.rstrip("\r\n").split("\t") for ln in f if ln.strip("\r\n")]
    cols = []
    for c, (t, w) in enumerate(descs):
        if t == AttrType.attrInteger:
            cols.append(np.array([int(r[c]) for r in rows], dtype=np.int32))
        else:
            a = np.zeros((len(rows), w), dtype=np.uint8)
            for i, r in enumerate(rows):
                b = r[c].encode("utf-8")
                if len(r[c]) > w or len(b) > w:
                    raise Exception("column value exceeds size limit")
                a[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
            cols.append(a)
    return names, descs, cols


class BatchInsert:
    """input/BatchInsert.java:21-140: `batchinsert DATAFILE DB CF NUMCOLUMNS`.  The data file is parsed on the host and
    the columns are loaded into a GPU-resident Columnarfile in one step (the reference inserts row by row into heap
    files; writing those pages stays in Java).  Together with `query ... FILESCAN` this is BASELINE config C1."""

    def insert(self, args: Sequence[str], echo: bool = True) -> list[str]:
        if len(args) < 4:
            raise Exception("Invalid number of attributes.")
        datafile, cfname = args[0], args[2]
        try:
            numcolumns = int(args[3])
        except ValueError:
            raise Exception("NUMCOULMNS is not integer.")
        names, descs, cols = parse_datafile(datafile, numcolumns)
        cf = Columnarfile(cfname, numcolumns, [AttrType(t) for t, _ in descs], [w for _, w in descs], names)
        cf.load_columns(cols)
        lines: list[str] = []
        _emit(lines, "Record count: " + str(cf.getTupleCnt()), echo)
        self.recordCount = cf.getTupleCnt()
        return lines


class Index:
    """input/Index.java:16-66: `index DB CF COL bitmap`"""

    def createIndex(self, args: Sequence[str]) -> None:
        _, cfname, colname, kind = args[:4]
        if kind.lower() != "bitmap":
            raise Exception("Only bitmap indexes are built on the GPU; B-tree indexes stay in Java")
        cf = Columnarfile(cfname)
        cf.createBitMapIndex(cf.colNameToIndex(colname))


class Query:
    """input/Query.java:35-155,248-297: `query DB CF [targets] {col,op,val} NUMBUF FILESCAN|COLUMNSCAN|BITMAP`"""

    def execute(self, args: Sequence[str], echo: bool = True) -> list[str]:
        if len(args) < 6:
            raise Exception("Invalid number of attributes.")
        cfname, targets, constraint, access = args[1], args[2], args[3], args[5]
        if not (targets.startswith("[") and targets.endswith("]")):
            raise Exception("[TARGETCOLUMNNAMES] format invalid.")
        if not (constraint.startswith("{") and constraint.endswith("}")):
            raise Exception("VALUECONSTRAINT format invalid.")
        if access.upper() not in ("FILESCAN", "COLUMNSCAN", "BTREE", "BITMAP"):
            raise Exception("access type invalid.")
        if access.upper() == "BTREE":
            raise Exception("BTREE access stays in Java (out of scope for the GPU path)")
        cf = Columnarfile(cfname)
        names = [t.strip() for t in targets[1:-1].split(",")]
        cols = [cf.colNameToIndex(n) for n in names]
        proj = [FldSpec(RelSpec(RelSpec.outer), c + 1) for c in cols]
        out_types = [cf.getAttributeTypes()[c].attrType for c in cols]
        parts = constraint[1:-1].strip().split(",")
        if len(parts) != 3:
            raise Exception("Invalid VALUECONSTRAINT elements")
        exprs, itypes, fnums, inames = build_cnf_condexpr("{(" + ",".join(parts) + ")}", cf)
        lines: list[str] = []
        _emit(lines, ", ".join(names), echo)
        count = 0
        if access.upper() == "BITMAP":                            # Query.executeBitmapScan (:248-297)
            if not cf.bitmapIndexExists(cf.colNameToIndex(parts[0].strip())):
                raise Exception("Bitmap index does not exist on column " + parts[0])
            it = ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(),
                                   len(cols), cols, proj, exprs, False)
        elif access.upper() == "COLUMNSCAN":                      # Query.executeColumnScan (:157-196)
            import copy
            colscan = copy.deepcopy(exprs)                        # buildQueryCondExprColscan: the column is field 1 of the
            for e in colscan:                                     # one-column predicate tuple
                if e is not None and e.type1.attrType == AttrType.attrSymbol:
                    e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), 1)
            it = ColumnarColumnScan(cf, cf.colNameToIndex(parts[0].strip()), len(cols), cols, proj, colscan)
        else:                                                     # executeFileScan (:121-155)
            it = ColumnarFileScan(cfname, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), len(cols), proj, exprs)
        while True:
            t = it.get_next()
            if t is None:
                break
            _emit(lines, _fmt(t, out_types), echo)
            count += 1
        it.close()
        _footer(lines, count, echo)
        self.resultCount = count
        return lines


class MultiIndexQuery:
    """input/MultiIndexQuery.java:99-136: `indexes_query DB CF [targets] CNF NUMBUF`"""

    def execute(self, args: Sequence[str], echo: bool = True) -> list[str]:
        cfname, targets, cnf = args[1], args[2], args[3]
        cf = Columnarfile(cfname)
        names = [t.strip() for t in targets[1:-1].split(",")]
        cols = [cf.colNameToIndex(n) for n in names]
        proj = [FldSpec(RelSpec(RelSpec.outer), c + 1) for c in cols]
        exprs, itypes, fnums, inames = build_cnf_condexpr(cnf, cf)
        it = ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(),
                               len(cols), cols, proj, exprs, False)
        out_types = [cf.getAttributeTypes()[c].attrType for c in cols]
        lines: list[str] = []
        _emit(lines, ", ".join(names), echo)
        count = 0
        while True:
            t = it.get_next()
            if t is None:
                break
            _emit(lines, _fmt(t, out_types), echo)
            count += 1
        it.close()
        _footer(lines, count, echo)
        self.resultCount = count
        return lines


class BitMapQuery:
    """input/BitMapQuery.java:49-305: `bmj DB OUTER INNER OUTERCNF INNERCNF JOINCNF [targets] NUMBUF`.
    The side filters run as bitmap CNF scans (getConstraintBitset :322-345), the join itself is K6."""

    def execute(self, args: Sequence[str], echo: bool = True) -> list[str]:
        if len(args) < 8:
            raise Exception("Invalid number of attributes.")
        outer_name, inner_name, ocnf, icnf, jcnf, targets = args[1], args[2], args[3], args[4], args[5], args[6]
        outer, inner = Columnarfile(outer_name), Columnarfile(inner_name)
        lines: list[str] = []
        osel = self._constraint_bitset(outer, ocnf)
        _emit(lines, "OuterConstraint Bitset After performing AND and ORs", echo)
        _emit(lines, repr(osel.getOutputPositions()), echo)
        isel = self._constraint_bitset(inner, icnf)
        _emit(lines, "InnerConstraint Bitset After performing AND and ORs", echo)
        _emit(lines, repr(isel.getOutputPositions()), echo)
        if not (targets.startswith("[") and targets.endswith("]")):
            raise Exception("[TARGETCOLUMNNAMES] format invalid.")
        names = [t.strip() for t in targets[1:-1].split(",")]
        proj, out_types = [], []
        for n in names:                                            # createProjectionsAndTuples (:113-185)
            rel, col = n.split(".")
            if rel == outer_name:
                proj.append((N.OPERAND_OUTER, outer.colNameToIndex(col)))
                out_types.append(outer.getAttributeTypes()[proj[-1][1]].attrType)
            else:
                proj.append((N.OPERAND_INNER, inner.colNameToIndex(col)))
                out_types.append(inner.getAttributeTypes()[proj[-1][1]].attrType)
        join_terms = []
        for ci, conj in enumerate(jcnf.split("^")):                # buildCNFJoinCondExprForFilling (:422-476)
            if not (conj.startswith("{") and conj.endswith("}")):
                raise Exception("Invalid query format")
            for dis in conj[1:-1].split("|"):
                parts = [p.strip() for p in dis[1:-1].strip().split(",")]
                if len(parts) != 3:
                    raise Exception("Invalid VALUECONSTRAINT elements")
                oc, ic = outer.colNameToIndex(parts[0]), inner.colNameToIndex(parts[2])
                if outer.getAttributeTypes()[oc].attrType != inner.getAttributeTypes()[ic].attrType:
                    raise Exception("Invalid JOIN COLUMN ATTR TYPE NOT MATCH.")
                if not inner.bitmapIndexExists(ic):                # the reference probes the inner column's bitmap index
                    raise Exception("Bitmap index does not exist on column " + parts[2])
                join_terms.append(Term(AttrOperator.findOperator(parts[1]).attrOperator, ("col", oc), ("icol", ic), ci))
        res = bitmap_join(outer.table, inner.table, join_terms, proj,
                          N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST, aggs=[(N.AGG_COUNT, 0)],
                          outer_sel=osel._result, inner_sel=isel._result)
        _emit(lines, ", ".join(names), echo)
        from .heap import Tuple
        for raw in res.tuples():
            t = Tuple(bytes(raw))
            t._adopt_header()
            _emit(lines, _fmt(t, out_types), echo)
        count = res.count
        res.close(); osel.close(); isel.close()
        _footer(lines, count, echo)
        self.resultCount = count
        return lines

    @staticmethod
    def _constraint_bitset(cf: Columnarfile, cnf: str) -> ColumnarIndexScan:
        exprs, itypes, fnums, inames = build_cnf_condexpr(cnf, cf)
        return ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), exprs)


def nlj_emission_order(outer_qualifying_positions, pair_outer, pair_inner, block_rows: int) -> np.ndarray:
    """Permutation that puts a pair list into the emission order of iterator/ColumnarNestedLoopJoins.java:157-207: the
    qualifying outer rows are taken block_rows at a time; per block every inner row in order, the block's outer rows
    inside.  Sort key = (outer block, inner position, outer position)."""
    po, pi = np.asarray(pair_outer, dtype=np.int64), np.asarray(pair_inner, dtype=np.int64)
    rank = np.searchsorted(np.asarray(outer_qualifying_positions, dtype=np.int64), po)
    return np.lexsort((po, pi, rank // max(int(block_rows), 1)))


class NljQuery:
    """input/NljQuery.java:33-330: `nlj DB OUTER INNER OUTERCONST INNERCONST JOINCONST OUTERACCESS INNERACCESS [targets]
    NUMBUF MEM`.  The reference runs a block nested-loop join over two column scans
    (iterator/ColumnarNestedLoopJoins.java:157-207); here both side constraints are GPU filter scans that leave
    selection bitmaps and the join is the same K6 as `bmj` (equi path or tiled theta kernel).  The pair list is then put
    into the reference's emission order: the qualifying outer rows are taken a block at a time -- (MEM - 1) pages of
    1024 / outer-tuple-size tuples, :118-121 -- and for every block the inner rows are walked in order, the block's
    outer rows inside (sort key: outer block, inner position, outer position), so the printed rows match the Java line
    by line."""

    def execute(self, args: Sequence[str], echo: bool = True, via_iterators: bool = False) -> list[str]:
        """via_iterators=True builds the two input scans and a ColumnarNestedLoopJoins over them exactly like
        NljQuery.java:139-176 (FILESCAN / COLUMNSCAN access); the default goes to the join kernels directly."""
        if len(args) < 11:
            raise Exception("Invalid number of attributes.")
        if via_iterators:
            return self._execute_via_iterators(args, echo)
        outer_name, inner_name, ocnf, icnf, jcnf, oacc, iacc, targets = args[1:9]
        for acc in (oacc, iacc):
            if acc.upper() not in ("FILESCAN", "COLUMNSCAN", "BTREE", "BITMAP"):
                raise Exception("access type invalid.")
            if acc.upper() == "BTREE":
                raise Exception("BTREE access stays in Java (out of scope for the GPU path)")
        if not (targets.startswith("[") and targets.endswith("]")):
            raise Exception("[TARGETCOLUMNNAMES] format invalid.")
        try:
            amt_of_memory = int(args[10])
        except ValueError:
            raise Exception("amt_of_memory is not integer.")
        if amt_of_memory < 2:
            raise Exception("amt_of_memory is not more than 1.")
        outer, inner = Columnarfile(outer_name), Columnarfile(inner_name)
        osel = self._constraint(outer, ocnf, oacc)
        isel = self._constraint(inner, icnf, iacc)
        names = [t.strip() for t in targets[1:-1].split(",")]
        proj, out_types = [], []
        for n in names:
            rel, col = n.split(".")
            cf, side = (outer, N.OPERAND_OUTER) if rel == outer_name else (inner, N.OPERAND_INNER)
            proj.append((side, cf.colNameToIndex(col)))
            out_types.append(cf.getAttributeTypes()[proj[-1][1]].attrType)
        join_terms = []
        for ci, conj in enumerate(jcnf.split("^")):
            if not (conj.startswith("{") and conj.endswith("}")):
                raise Exception("Invalid query format")
            for dis in conj[1:-1].split("|"):
                parts = [p.strip() for p in dis[1:-1].strip().split(",")]
                if len(parts) != 3:
                    raise Exception("Invalid VALUECONSTRAINT elements")
                oc, ic = outer.colNameToIndex(parts[0]), inner.colNameToIndex(parts[2])
                if outer.getAttributeTypes()[oc].attrType != inner.getAttributeTypes()[ic].attrType:
                    raise Exception("Invalid JOIN COLUMN ATTR TYPE NOT MATCH.")
                join_terms.append(Term(AttrOperator.findOperator(parts[1]).attrOperator, ("col", oc), ("icol", ic), ci))
        res = bitmap_join(outer.table, inner.table, join_terms, proj,
                          N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST, aggs=[(N.AGG_COUNT, 0)],
                          outer_sel=osel, inner_sel=isel)
        # the outer iterator's tuple (NljQuery.java:84-106,446-470): target columns of the outer file, the join columns,
        # and -- unless the access is FILESCAN -- the constraint columns of every conjunct after the first (a TreeSet)
        tcols = {c for side, c in proj if side == N.OPERAND_OUTER} | {t.lhs[1] for t in join_terms}
        if oacc.upper() != "FILESCAN":
            for conj in ocnf.split("^")[1:]:
                tcols |= {outer.colNameToIndex(d[1:-1].split(",")[0].strip()) for d in conj[1:-1].split("|")}
        sizes, types = outer.getAttrSizes(), outer.getAttributeTypes()
        tuple_size = (len(tcols) + 2) * 2 + sum(sizes[c] + 2 if types[c].attrType == AttrType.attrString else 4 for c in tcols)
        block_rows = max(1, (amt_of_memory - 1) * (1024 // tuple_size))
        order = nlj_emission_order(osel.positions(), res.positions(), res.positions2(), block_rows)
        lines: list[str] = []
        _emit(lines, ", ".join(names), echo)
        from .heap import Tuple
        tuples = res.tuples()
        for k in order:
            t = Tuple(bytes(tuples[k]))
            t._adopt_header()
            _emit(lines, _fmt(t, out_types), echo)
        count = res.count
        res.close(); osel.close(); isel.close()
        _footer(lines, count, echo)
        self.resultCount = count
        return lines

    def _execute_via_iterators(self, args: Sequence[str], echo: bool) -> list[str]:
        from .iterator import ColumnarColumnsScan, ColumnarNestedLoopJoins
        outer_name, inner_name, ocnf, icnf, jcnf, oacc, iacc, targets = args[1:9]
        amt_of_memory = int(args[10])
        files = {"o": Columnarfile(outer_name), "i": Columnarfile(inner_name)}
        names = [t.strip() for t in targets[1:-1].split(",")]
        tcols = {"o": set(), "i": set()}                            # the TreeSets of NljQuery.java:84-106
        for n in names:
            rel, col = n.split(".")
            side = "o" if rel == outer_name else "i"
            tcols[side].add(files[side].colNameToIndex(col))
        for side, cnf, acc in (("o", ocnf, oacc), ("i", icnf, iacc)):   # findConsTargetCols (:446-470)
            if acc.upper() not in ("FILESCAN", "COLUMNSCAN"):
                raise Exception("via_iterators serves FILESCAN and COLUMNSCAN access")
            if acc.upper() != "FILESCAN":
                for conj in cnf.split("^")[1:]:
                    tcols[side] |= {files[side].colNameToIndex(d[1:-1].split(",")[0].strip()) for d in conj[1:-1].split("|")}
        join = []
        for ci, conj in enumerate(jcnf.split("^")):                 # findJoinTargetCols + buildCNFJoinCondExpr
            for dis in conj[1:-1].split("|"):
                a, op, b = [p.strip() for p in dis[1:-1].strip().split(",")]
                oc, ic = files["o"].colNameToIndex(a), files["i"].colNameToIndex(b)
                tcols["o"].add(oc); tcols["i"].add(ic)
                join.append((ci, op, oc, ic))
        order = {s: sorted(tcols[s]) for s in "oi"}
        field = {s: {c: k + 1 for k, c in enumerate(order[s])} for s in "oi"}      # findFieldOffset

        def iterator_of(side, cnf, acc):                            # getIterator (:259-300)
            cf = files[side]
            exprs, _, _, _ = build_cnf_condexpr(cnf, cf)
            proj = [FldSpec(RelSpec(RelSpec.outer), c + 1) for c in order[side]]
            if acc.upper() == "FILESCAN":
                return ColumnarFileScan(cf._fileName, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), len(proj), proj, exprs)
            used = sorted({t.operand1.symbol.offset - 1 for e in exprs if e is not None for t in _chain(e)})
            for e in exprs:
                if e is not None:
                    for t in _chain(e):
                        t.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), used.index(t.operand1.symbol.offset - 1) + 1)
            return ColumnarColumnsScan(cf, used, len(proj), order[side], proj, exprs)

        def _chain(e):
            while e is not None:
                yield e
                e = e.next

        outer_it, inner_it = iterator_of("o", ocnf, oacc), iterator_of("i", icnf, iacc)
        join_filter, by_conj = [], {}
        for ci, op, oc, ic in join:
            e = CondExpr()
            e.op = AttrOperator.findOperator(op)
            e.type1 = e.type2 = AttrType(AttrType.attrSymbol)
            e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), field["o"][oc])
            e.operand2.symbol = FldSpec(RelSpec(RelSpec.innerRel), field["i"][ic])
            if ci in by_conj:
                tail = by_conj[ci]
                while tail.next is not None:
                    tail = tail.next
                tail.next = e
            else:
                by_conj[ci] = e
                join_filter.append(e)
        join_filter.append(None)
        proj_list, out_types = [], []
        for n in names:
            rel, col = n.split(".")
            side = "o" if rel == outer_name else "i"
            c = files[side].colNameToIndex(col)
            proj_list.append(FldSpec(RelSpec(RelSpec.outer if side == "o" else RelSpec.innerRel), field[side][c]))
            out_types.append(files[side].getAttributeTypes()[c].attrType)

        def tuple_types(side):
            cf = files[side]
            types = [cf.getAttributeTypes()[c] for c in order[side]]
            return types, [cf.getAttrSizes()[c] for c in order[side] if cf.getAttributeTypes()[c].attrType == AttrType.attrString]

        (in1, s1), (in2, s2) = tuple_types("o"), tuple_types("i")
        nlj = ColumnarNestedLoopJoins(files["o"], files["i"], in1, len(in1), s1, in2, len(in2), s2, outer_it, inner_it,
                                      None, None, join_filter, proj_list, len(proj_list), amt_of_memory)
        lines: list[str] = []
        _emit(lines, ", ".join(names), echo)
        count = 0
        while True:
            t = nlj.get_next()
            if t is None:
                break
            _emit(lines, _fmt(t, out_types), echo)
            count += 1
        nlj.close()
        _footer(lines, count, echo)
        self.resultCount = count
        return lines

    @staticmethod
    def _constraint(cf: Columnarfile, cnf: str, access: str):
        """The side's qualifying rows as a selection bitmap: a filter scan (FILESCAN / COLUMNSCAN) or the bitmap
        indexes (BITMAP)."""
        exprs, itypes, fnums, inames = build_cnf_condexpr(cnf, cf)
        if access.upper() == "BITMAP":
            scan = ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), exprs)
            scan.getOutputPositions()
            res, scan._result = scan._result, None                 # the selection outlives the scan object
            return res
        from .iterator import flatten_condexpr
        return cf.table.scan(flatten_condexpr(exprs), want=N.WANT_BITMAP | N.WANT_POSITIONS | N.WANT_HOST)


class ColumnarSort:
    """input/ColumnarSort.java:73-400: `sort DB CF [sort columns] [projected columns] ASC|DSC NUMBUF SORTBUF`
    (SURVEY.md 8f rank 4).  The reference runs an external merge sort over (keys, position) records; here the row ids
    are radix-sorted on the GPU (mbc_sort) and the projected fields gathered in that order.  Same printed lines
    `<projected values> :<position>`, equal keys in the order the Java's merge structure leaves them (replayed on the
    host for results up to REFERENCE_TIE_ORDER_MAX_ROWS rows; ascending position beyond that)."""

    def execute(self, args: Sequence[str], echo: bool = True) -> list[str]:
        if len(args) < 7:
            raise Exception("Invalid number of attributes.")
        cfname, sort_cols, proj_cols, order = args[1], args[2], args[3], args[4]
        if not (sort_cols.startswith("[") and sort_cols.endswith("]")):
            raise Exception("[TARGETCOLUMNNAMES] format invalid.")
        if not (proj_cols.startswith("[") and proj_cols.endswith("]")):
            raise Exception("[PROJECTIONCOLUMNNAMES] format invalid.")
        if order not in ("ASC", "DSC"):
            raise Exception("This sorting order is not supported")
        try:
            numbuf, sortbuf = int(args[5]), int(args[6])
        except ValueError:
            raise Exception("NUMBUF is not integer.")
        if numbuf < 1:
            raise Exception("NUMBUF is not more than 1.")
        lines: list[str] = []
        if sortbuf < 3:
            _emit(lines, "NUMBUF_SORT is less than 3. External Sort Merge needs minimum 3 pages for the operation", echo)
            return lines
        cf = Columnarfile(cfname)
        keys = [cf.colNameToIndex(n.strip()) for n in sort_cols[1:-1].split(",")]
        proj = [cf.colNameToIndex(n.strip()) for n in proj_cols[1:-1].split(",")]
        types = [cf.getAttributeTypes()[c].attrType for c in proj]
        res = cf.table.sort(keys, descending=(order == "DSC"), proj=keys + proj,
                            want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_HOST)
        _emit(lines, "SORTED COLUMNS", echo)
        pos = res.positions()
        key_cols = [np.asarray(res.column(i)) for i in range(len(keys))]
        cols = [res.column(len(keys) + i) for i in range(len(proj))]
        emit = range(res.count)
        if 0 < res.count <= self.REFERENCE_TIE_ORDER_MAX_ROWS:
            emit = self._reference_order(cf, keys, key_cols, pos, sortbuf)
        for k in emit:                                              # printRecordsByPages: values, blank separated, then :position
            vals = [str(int(cols[f][k])) if t == AttrType.attrInteger else repr(float(cols[f][k])) if t == AttrType.attrReal
                    else bytes(cols[f][k]).rstrip(b"\0").decode("utf-8") for f, t in enumerate(types)]
            _emit(lines, " ".join(vals) + " :" + str(int(pos[k])), echo)
        _emit(lines, str(res.count), echo)
        self.resultCount = res.count
        res.close()
        return lines

    # The GPU sort leaves equal keys in ascending position.  The Java's external merge sort leaves them in an order that
    # follows from its page and run structure (ColumnarSort.java:206-352); for tables up to this many rows the mirror
    # replays that structure on the host over the key groups of the GPU result, so the lines come out exactly as the
    # reference prints them.  Larger results keep the GPU order (a valid sort either way).
    REFERENCE_TIE_ORDER_MAX_ROWS = 200_000

    @staticmethod
    def _reference_order(cf: Columnarfile, keys: Sequence[int], key_cols, pos, sort_buffers: int) -> list:
        """Indices into the GPU-sorted result, in the order the reference emits the rows."""
        n = len(pos)
        group = np.zeros(n, dtype=np.int64)                        # key rank of every row of the sorted result
        if n > 1:
            change = np.zeros(n - 1, dtype=bool)
            for kc in key_cols:
                a = np.asarray(kc).reshape(n, -1)
                change |= (a[1:] != a[:-1]).any(axis=1)
            group[1:] = np.cumsum(change)
        by_pos = np.argsort(pos, kind="stable")                    # record r of the sort heap file = r-th live row by position
        sizes, types = cf.getAttrSizes(), cf.getAttributeTypes()
        rec = sum(sizes[c] + 2 if types[c].attrType == AttrType.attrString else 4 for c in keys) + 4
        records = _external_sort_replay(group[by_pos].tolist(), 1004 // (rec + 4), sort_buffers - 1)
        return [int(by_pos[r]) for r in records]


def _external_sort_replay(key: list, per_page: int, nbuf: int) -> list:
    """The pass structure of input/ColumnarSort.java:206-352 on record numbers 0..n-1 with integer keys: P records per
    heap page, pass 0 reads B pages round-robin and sorts them stably, every later pass merges B runs of B^i pages taking
    the smallest head, the lowest run among equal heads.  Returns the record numbers in emission order."""
    import heapq
    import math
    n = len(key)
    npages = (n + per_page - 1) // per_page
    passes = int(math.ceil(math.log(npages) / math.log(nbuf))) if npages > 1 and nbuf > 1 else 0
    cur = list(range(n))
    for i in range(passes):
        pages = [cur[s:s + per_page] for s in range(0, n, per_page)]
        run = nbuf ** i
        runs = [pages[s:s + run] for s in range(0, npages, run)]
        out = []
        for k in range(0, len(runs), nbuf):
            grp = runs[k:k + nbuf]
            if i == 0:
                heads = [r[0] for r in grp]
                lst = [h[s] for s in range(max(len(h) for h in heads)) for h in heads if s < len(h)]
                lst.sort(key=key.__getitem__)                      # Collections.sort: stable
                out += lst
            else:
                seqs = [[(key[r], j, q, r) for q, r in enumerate(x for pg in rn for x in pg)] for j, rn in enumerate(grp)]
                out += [t[3] for t in heapq.merge(*seqs)]
        cur = out
    return cur


class DeleteQuery:
    """input/DeleteQuery.java:27-205: `delete_query DB CF {col,op,val} NUMBUF FILESCAN|COLUMNSCAN|BITMAP md|pd`.
    The qualifying TIDs come from the tid-only constructors of ColumnarFileScan / ColumnarColumnScan / ColumnIndexScan
    (`get_next_tid`), every one is marked deleted (Columnarfile.markTupleDeleted, :812) and every later scan skips those
    rows.  `pd` (purge: rewrite the heap files without the rows) is a mutation of the storage engine and stays in Java."""

    def execute(self, args: Sequence[str], echo: bool = True) -> list[str]:
        if len(args) < 6:
            raise Exception("Invalid number of attributes.")
        cfname, constraint, access, delete_type = args[1], args[2], args[4], args[5]
        if not (constraint.startswith("{") and constraint.endswith("}")):
            raise Exception("VALUECONSTRAINT format invalid.")
        try:
            if int(args[3]) < 1:
                raise Exception("NUMBUF is not more than 1.")
        except ValueError:
            raise Exception("NUMBUF is not integer.")
        if access.upper() not in ("FILESCAN", "COLUMNSCAN", "BTREE", "BITMAP"):
            raise Exception("access type invalid.")
        if delete_type.lower() not in ("md", "pd"):
            raise Exception("delete type invalid.")
        if access.upper() == "BTREE":
            raise Exception("BTREE access stays in Java (out of scope for the GPU path)")
        if delete_type.lower() == "pd":
            raise Exception("purge (pd) rewrites the heap files: stays in Java; use md")
        cf = Columnarfile(cfname)
        parts = constraint[1:-1].strip().split(",")
        if len(parts) != 3:
            raise Exception("Invalid VALUECONSTRAINT elements")
        exprs, itypes, fnums, inames = build_cnf_condexpr("{(" + ",".join(parts) + ")}", cf)
        col = cf.colNameToIndex(parts[0].strip())
        if access.upper() == "FILESCAN":                           # executeFileScan (:119-133)
            it = ColumnarFileScan(cfname, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), exprs)
        elif access.upper() == "COLUMNSCAN":                       # executeColumnScan (:135-153)
            import copy
            colscan = copy.deepcopy(exprs)
            for e in colscan:
                if e is not None and e.type1.attrType == AttrType.attrSymbol:
                    e.operand1.symbol = FldSpec(RelSpec(RelSpec.outer), 1)
            it = ColumnarColumnScan(cf, col, colscan)
        else:                                                      # executeBitmapScan (:181-205)
            if not cf.bitmapIndexExists(col):
                raise Exception("Bitmap index does not exist on column " + parts[0])
            it = ColumnarIndexScan(cf, fnums, itypes, inames, cf.getAttributeTypes(), cf.getStringSizes(), cf.getFieldCount(), exprs)
        tids = []
        if isinstance(it, ColumnarIndexScan):
            tids = [int(p) for p in it.getOutputPositions().positions()]
        else:
            while True:
                tid = it.get_next_tid()
                if tid is None:
                    break
                tids.append(tid)
        it.close()
        self.deletedCount = cf.markTuplesDeleted(tids)
        lines: list[str] = []
        _emit(lines, "=======================EXTRA METAINFO===============================", echo)
        _emit(lines, str(cf.getTupleCnt()), echo)
        _emit(lines, repr(cf.getMarkedDeleted().getBitSet()), echo)  # Columnarfile.printDeleteBitset (:573)
        return lines
