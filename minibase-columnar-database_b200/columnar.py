"""Mirror of columnar.Columnarfile's read side (minijava/src/columnar/Columnarfile.java): open by name,
schema accessors, bitmap index build / lookup, markedDeleted, tuple and column scans.  The columns live in
HBM as one contiguous array each (the reference keeps one heapfile per column, :329-337)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .bitmap import BitMapFile, BitSet
from .engine import Table
from .global_ import AttrType, IntegerValue, StringValue, SystemDefs, TID, ValueClass
from .heap import Tuple


class Columnarfile:
    """Columnarfile(name)                                   open an existing file (:239-359): looked up in the
                                                            current SystemDefs, ingested from its DB image (K1) the
                                                            first time
       Columnarfile(name, numColumns, types, sizes, names)  create (:43-140); rows come through load_columns /
                                                            generate (the reference's row-at-a-time insertTuple stays in Java)"""

    def __init__(self, name: str, numColumns: Optional[int] = None, attrTypes: Optional[Sequence[AttrType]] = None,
                 attrSizes: Optional[Sequence[int]] = None, attrNames: Optional[Sequence[str]] = None):
        sd = SystemDefs.current()
        self._fileName = name
        if numColumns is None:
            existing = sd.files.get(name)
            if existing is not None:
                self.__dict__ = existing.__dict__                # same open file (the reference re-reads .hdr)
                return
            if sd.db_bytes is None:
                raise Exception("Columnar File does not exist.")   # Columnarfile.java:253
            try:
                self.table = sd.ctx.ingest_dbfile(sd.db_bytes, name)
            except N.MbcError as e:
                raise Exception(e.message)
            from .dbfile import read_header
            hdr = read_header(sd.db_bytes, name)
            self._from_image = True
            self.attrNames = hdr["colnames"]
            self._init_schema([(t, w) for t, w in self.table.coldescs])
            if hdr["deleted_bytes"]:                               # the device copy was set by the ingest; mirror it
                self._deleted.bitSet = BitSet(np.frombuffer(hdr["deleted_bytes"], dtype=np.uint64).copy())
            for c, flag in enumerate(hdr["bitmapExist"]):          # Columnarfile.java:288-323: the catalogued indexes are
                if flag == 1:                                      # available again; the device copy is rebuilt from the column
                    self.table.bitmap_build(c)                     # (K3 takes less time than decoding the BM pages would)
        else:
            if any(len(n) > 15 for n in attrNames):
                raise Exception("Attribute name too long.")        # Columnarfile.java:68-70 (MAXATTRNAME)
            self.attrNames = list(attrNames)
            self.table = None
            self._pending = [(int(t.attrType), int(s)) for t, s in zip(attrTypes, attrSizes)]
            self._init_schema(self._pending)
        sd.files[name] = self

    def _init_schema(self, coldescs) -> None:
        self.numColumns = len(coldescs)
        self.attrTypes = [AttrType(t) for t, _ in coldescs]
        self.attrSizes = [w for _, w in coldescs]
        self.stringSizes = [w for t, w in coldescs if t == AttrType.attrString]
        self._deleted = BitMapFile(BitSet(), on_change=self._push_deleted)

    # ---- bulk load (replaces BatchInsert's per-row insertTuple loop for the GPU-resident copy) ----------
    def load_columns(self, columns: Sequence[np.ndarray]) -> None:
        descs = list(zip([t.attrType for t in self.attrTypes], self.attrSizes))
        nrows = columns[0].size // descs[0][1] if descs[0][0] == AttrType.attrString else np.asarray(columns[0]).size
        self.table = SystemDefs.current().ctx.create_table(descs, nrows)
        for c, col in enumerate(columns):
            self.table.load_column(c, col)

    def close(self) -> None:
        if getattr(self, "table", None) is not None:
            self.table.close()
            self.table = None
        SystemDefs.current().files.pop(self._fileName, None)

    # ---- schema accessors (Columnarfile.java:1027-1100) ------------------------------------------------------
    def getFieldCount(self) -> int:
        return self.numColumns

    def getAttributeTypes(self):
        return self.attrTypes

    def getAttributeType(self, col: int) -> AttrType:
        return self.attrTypes[col]

    def getAttrSizes(self):
        return self.attrSizes

    def getStringSizes(self):
        return self.stringSizes

    def getAttrNames(self):
        return self.attrNames

    def getTupleCnt(self) -> int:
        return self.table.nrows

    def colNameToIndex(self, name: str) -> int:
        if name in self.attrNames:
            return self.attrNames.index(name)
        raise Exception(f"Column Name '{name}' Invalid.")        # Columnarfile.java:1047

    def indexToColName(self, index: int) -> str:
        if index < len(self.attrNames):
            return self.attrNames[index]
        raise Exception(f"Column Index '{index}' out of bound.")

    # ---- bitmap indexes ------------------------------------------------------------------------------------------
    def createBitMapIndex(self, columnNo: int) -> bool:
        """Columnarfile.java:698-753 -> K3 on the device.  When the file was opened from a DB image, the index is also
        PERSISTED into that image in the reference's own format (one `<cf>.bm.<col>.<value>` BMIndexPage chain per value,
        the "<col>.<value>" catalogue records and bitmapExist in `<cf>.hdr`; dbfile.persist_bitmap_index), so the unmodified
        Java opens it with BitMapFile(String); SystemDefs.flush() writes the image back to the DB file."""
        self.table.bitmap_build(columnNo)
        sd = SystemDefs.current()
        if getattr(self, "_from_image", False) and sd.db_bytes is not None:
            from .dbfile import persist_bitmap_index
            vals = self.table.bitmap_values(columnNo)
            if self.attrTypes[columnNo].attrType == AttrType.attrString:
                values = [bytes(v).rstrip(b"\0").decode("utf-8") for v in vals]
            else:
                values = [int(v) for v in vals]
            bitsets = [self.table.bitmap_get(columnNo, v).view(np.uint8).tobytes() for v in values]
            sd.db_bytes = persist_bitmap_index(sd.db_bytes, self._fileName, columnNo, values, bitsets)
        return True

    def bitmapIndexExists(self, colNo: int) -> bool:
        return self.table.bitmap_exists(colNo)

    def getBitmapValues(self, columnNo: int) -> set:
        """Columnarfile.java:1096: the distinct indexed values (a HashSet in the reference: order unspecified)."""
        vals = self.table.bitmap_values(columnNo)
        if self.attrTypes[columnNo].attrType == AttrType.attrString:
            return {bytes(v).rstrip(b"\0").decode("utf-8") for v in vals}
        return {int(v) for v in vals}

    def getBitmapIndex(self, columnNo: int, value) -> BitMapFile:
        """Columnarfile.java:1103-1127: the value's BitMapFile; an empty one when the value was never indexed."""
        v = value.getValue() if isinstance(value, ValueClass) else value
        return BitMapFile(BitSet(self.table.bitmap_get(columnNo, v)))

    # ---- deleted rows ----------------------------------------------------------------------------------------------------
    def getMarkedDeleted(self) -> BitMapFile:
        return self._deleted

    def _push_deleted(self, bits: BitSet) -> None:
        self.table.set_deleted(bits.words if bits.words.size else np.zeros(1, dtype=np.uint64))

    def markTupleDeleted(self, position) -> bool:
        """Columnarfile.java:812 markTupleDeleted(TID) (a bare position is accepted too): only the GPU-resident deleted
        bitmap is maintained here."""
        self._deleted.set(position.position if isinstance(position, TID) else int(position))
        return True

    def markTuplesDeleted(self, positions) -> int:
        """The delete drivers' loop `while ((tid = scan.get_next_tid()) != null) markTupleDeleted(tid)` in one step: all
        bits are set on the host BitSet, the device copy is refreshed once."""
        bits = self._deleted.getBitSet()
        n = 0
        for p in positions:
            bits.set(p.position if isinstance(p, TID) else int(p))
            n += 1
        self._push_deleted(bits)
        return n

    # ---- scans ---------------------------------------------------------------------------------------------------------------
    def openTupleScan(self) -> "TupleScan":
        return TupleScan(self)


class TupleScan:
    """columnar/TupleScan.java:14-100: every live row, all columns, in position order."""

    def __init__(self, columnarFile: Columnarfile):
        from .iterator import _ResultCursor
        self.columnarFile = columnarFile
        n = columnarFile.numColumns
        self._result = columnarFile.table.scan([], proj=list(range(n)), want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_HOST)
        self._positions = self._result.positions()
        self._cols = [self._result.column(i) for i in range(n)]
        self._i = 0

    def getNext(self, tid: TID) -> Optional[Tuple]:
        if self._i >= len(self._positions):
            return None
        k = self._i
        self._i += 1
        cf = self.columnarFile
        t = Tuple()                                              # a fresh zeroed tuple per row (TupleScan.java:57)
        t.setHdr(cf.numColumns, cf.attrTypes, cf.stringSizes)
        for c, at in enumerate(cf.attrTypes):
            if at.attrType == AttrType.attrInteger:
                t.setIntFld(c + 1, int(self._cols[c][k]))
            elif at.attrType == AttrType.attrReal:
                t.setFloFld(c + 1, float(self._cols[c][k]))
            else:
                t.setStrFld(c + 1, bytes(self._cols[c][k]).rstrip(b"\0").decode("utf-8"))
        tid.setPosition(int(self._positions[k]))
        return t

    def closetuplescan(self) -> None:
        if self._result is not None:
            self._result.close()
            self._result = None
