"""Mirror of the reference's ``global`` package for the columnar scan path (``global`` is a Python keyword,
hence the trailing underscore).  Same class names, constants and method meaning as
minijava/src/global/*.java; file:line references point into /root/reference/minijava/src.
"""
from __future__ import annotations

from typing import Optional


class AttrType:
    """global/AttrType.java:10-14"""
    attrString, attrInteger, attrReal, attrSymbol, attrNull = 0, 1, 2, 3, 4

    def __init__(self, attrType: int):
        self.attrType = attrType

    def __repr__(self):
        return {0: "attrString", 1: "attrInteger", 2: "attrReal", 3: "attrSymbol", 4: "attrNull"}.get(
            self.attrType, f"Unexpected AttrType {self.attrType}")

    toString = __repr__


class AttrOperator:
    """global/AttrOperator.java:10-18,66-102"""
    aopEQ, aopLT, aopGT, aopNE, aopLE, aopGE, aopNOT, aopNOP, opRANGE = range(9)
    _NAMES = ["aopEQ", "aopLT", "aopGT", "aopNE", "aopLE", "aopGE", "aopNOT", "aopNOP", "opRANGE"]

    def __init__(self, attrOperator: int):
        self.attrOperator = attrOperator

    def toString(self) -> str:
        if 0 <= self.attrOperator < 9:
            return self._NAMES[self.attrOperator]
        return f"Unexpected AttrOperator {self.attrOperator}"

    __repr__ = toString

    @staticmethod
    def findOperator(op: str) -> "AttrOperator":                      # AttrOperator.java:66-83
        table = {"=": 0, "<": 1, ">": 2, "!=": 3, ">=": 5, "<=": 4}
        if op not in table:
            raise Exception("unsupported or invalid operator")
        return AttrOperator(table[op])

    @staticmethod
    def getOppositeOperator(op: str) -> "AttrOperator":               # AttrOperator.java:85-102
        table = {"=": 0, "<": 2, ">": 1, "!=": 3, ">=": 4, "<=": 5}
        if op not in table:
            raise Exception("unsupported or invalid operator")
        return AttrOperator(table[op])


class IndexType:
    """global/IndexType.java"""
    None_, B_Index, Hash, Bitmap = 0, 1, 2, 3

    def __init__(self, indexType: int):
        self.indexType = indexType

    def toString(self) -> str:
        return {0: "None", 1: "B_Index", 2: "Hash", 3: "Bitmap"}.get(self.indexType, f"Unexpected IndexType {self.indexType}")

    __repr__ = toString


class PageId:
    def __init__(self, pid: int = -1):
        self.pid = pid


class RID:
    """global/RID.java:10"""

    def __init__(self, pageNo: Optional[PageId] = None, slotNo: int = 0):
        self.pageNo = pageNo or PageId()
        self.slotNo = slotNo


class TID:
    """global/TID.java:8-29: a tuple id = the row's position plus one RID per column.  On the GPU path only the
    position is materialised (it is what every caller on the path reads); recordIDs stays None."""

    def __init__(self, numRIDs: int, position: int = -1, recordIDs=None):
        self.numRIDs = numRIDs
        self.position = position
        self.recordIDs = recordIDs

    def setPosition(self, position: int) -> None:
        self.position = position

    def copyTid(self, tid: "TID") -> None:
        self.numRIDs, self.position, self.recordIDs = tid.numRIDs, tid.position, tid.recordIDs

    def __eq__(self, other):
        return isinstance(other, TID) and (self.numRIDs, self.position) == (other.numRIDs, other.position)

    def __repr__(self):
        return f"TID(numRIDs={self.numRIDs}, position={self.position})"


class ValueClass:
    def getValue(self):
        return self.value


class IntegerValue(ValueClass):
    def __init__(self, value: int):
        self.value = int(value)


class StringValue(ValueClass):
    def __init__(self, value: str):
        self.value = value


class SystemDefs:
    """global/SystemDefs.java:7-96.  The reference's constructor opens the DB file and builds the buffer pool;
    here it opens a GPU context and (optionally) reads the DB file image that Columnarfile(name) ingests
    from.  JavabaseDB / JavabaseBM keep their names as class attributes so drivers written against the
    reference (``new SystemDefs(db, 0, numbuf, null)`` then ``new Columnarfile(name)``) keep working."""

    JavabaseDB = None       # the current SystemDefs instance (stands in for the DB singleton)
    JavabaseBM = None       # device HBM is the pool: this is the engine Context

    def __init__(self, dbname: Optional[str] = None, num_pgs: int = 0, bufpoolsize: int = 0,
                 replacement_policy: Optional[str] = None, device: int = 0):
        from .engine import Context
        self.dbname = dbname
        self.db_bytes = None
        if dbname is not None and num_pgs == 0:
            with open(dbname, "rb") as f:                     # existing database (num_pgs == 0 -> openDB)
                self.db_bytes = f.read()
        prev = SystemDefs.JavabaseDB
        self.ctx = prev.ctx if prev is not None and prev.ctx.device == device and prev.ctx._h else Context(device)
        self.files = {} if prev is None or prev.ctx is not self.ctx else prev.files      # name -> Columnarfile
        SystemDefs.JavabaseDB = self
        SystemDefs.JavabaseBM = self.ctx

    def flush(self) -> None:
        """Write the (possibly edited: persisted bitmap indexes) DB image back to the DB file -- the reference's buffer
        manager does this page by page on flushAllPages (bufmgr/BufMgr.java)."""
        if self.dbname is not None and self.db_bytes is not None:
            with open(self.dbname, "wb") as f:
                f.write(self.db_bytes)

    @classmethod
    def current(cls) -> "SystemDefs":
        if cls.JavabaseDB is None:
            cls()
        return cls.JavabaseDB

    @classmethod
    def shutdown(cls) -> None:
        if cls.JavabaseDB is not None:
            for f in list(cls.JavabaseDB.files.values()):
                f.close()
            cls.JavabaseDB.ctx.close()
        cls.JavabaseDB = None
        cls.JavabaseBM = None
