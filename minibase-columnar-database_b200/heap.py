"""Mirror of heap.Tuple (minijava/src/heap/Tuple.java): the byte-level tuple the scan operators return.

Wire format (Tuple.java:369-440, global/Convert.java:163-275), all big-endian:
    [fldCnt:short][fldOffset[0..n]:short x (n+1)] then the fields; fldOffset[0] = (n+2)*2;
    int/real = 4 bytes; string slot = strSize+2 bytes = [len:short][modified UTF-8][padding].
"""
from __future__ import annotations

import struct
from typing import Optional, Sequence

from .global_ import AttrType


class FieldNumberOutOfBoundException(Exception):
    pass


class InvalidTypeException(Exception):
    pass


class InvalidTupleSizeException(Exception):
    pass


class Tuple:
    max_size = 1024                                   # GlobalConst.MINIBASE_PAGESIZE

    def __init__(self, atuple: Optional[bytes] = None, offset: int = 0, length: Optional[int] = None):
        if atuple is None:
            self.data = bytearray(self.max_size)      # Tuple(): a fresh zeroed page-sized buffer
            self.tuple_offset = 0
            self.tuple_length = self.max_size
        else:
            self.data = atuple if isinstance(atuple, bytearray) else bytearray(atuple)
            self.tuple_offset = offset
            self.tuple_length = len(atuple) - offset if length is None else length
        self.fldCnt = 0
        self.fldOffset: list[int] = []

    # ---- header --------------------------------------------------------------------------------
    def setHdr(self, numFlds: int, types: Sequence[AttrType], strSizes: Optional[Sequence[int]]) -> None:
        """Tuple.java:369-440"""
        if (numFlds + 2) * 2 > self.max_size:
            raise InvalidTupleSizeException("TUPLE: TUPLE_TOOBIG_ERROR")
        self.fldCnt = numFlds
        off = (numFlds + 2) * 2 + self.tuple_offset
        self.fldOffset = [off]
        sc = 0
        for i in range(numFlds):
            t = types[i].attrType
            if t in (AttrType.attrInteger, AttrType.attrReal):
                incr = 4
            elif t == AttrType.attrString:
                incr = strSizes[sc] + 2
                sc += 1
            else:
                raise InvalidTypeException("TUPLE: TUPLE_TYPE_ERROR")
            off += incr
            self.fldOffset.append(off)
        struct.pack_into(">h", self.data, self.tuple_offset, numFlds)
        for i, o in enumerate(self.fldOffset):
            struct.pack_into(">h", self.data, self.tuple_offset + 2 + 2 * i, o)
        self.tuple_length = self.fldOffset[numFlds] - self.tuple_offset
        if self.tuple_length > self.max_size:
            raise InvalidTupleSizeException("TUPLE: TUPLE_TOOBIG_ERROR")

    def _adopt_header(self) -> None:
        """Read fldCnt / fldOffset back from the bytes (what Tuple(byte[],off,len)+setHdr gives the reference)."""
        n = struct.unpack_from(">h", self.data, self.tuple_offset)[0]
        self.fldCnt = n
        self.fldOffset = [struct.unpack_from(">h", self.data, self.tuple_offset + 2 + 2 * i)[0] for i in range(n + 1)]

    def _check(self, fldNo: int) -> None:
        if not (0 < fldNo <= self.fldCnt):
            raise FieldNumberOutOfBoundException("TUPLE:TUPLE_FLDNO_OUT_OF_BOUND")

    # ---- getters (1-based field numbers) -----------------------------------------------------------
    def getIntFld(self, fldNo: int) -> int:
        self._check(fldNo)
        return struct.unpack_from(">i", self.data, self.fldOffset[fldNo - 1])[0]

    def getFloFld(self, fldNo: int) -> float:
        self._check(fldNo)
        return struct.unpack_from(">f", self.data, self.fldOffset[fldNo - 1])[0]

    def getStrFld(self, fldNo: int) -> str:
        self._check(fldNo)
        o = self.fldOffset[fldNo - 1]
        ln = struct.unpack_from(">H", self.data, o)[0]
        return bytes(self.data[o + 2:o + 2 + ln]).decode("utf-8")      # ASCII/BMP contract: modified UTF-8 == UTF-8

    # ---- setters -------------------------------------------------------------------------------------
    def setIntFld(self, fldNo: int, val: int) -> "Tuple":
        self._check(fldNo)
        struct.pack_into(">i", self.data, self.fldOffset[fldNo - 1], val)
        return self

    def setFloFld(self, fldNo: int, val: float) -> "Tuple":
        self._check(fldNo)
        struct.pack_into(">f", self.data, self.fldOffset[fldNo - 1], val)
        return self

    def setStrFld(self, fldNo: int, val: str) -> "Tuple":
        """Convert.setStrValue (:254-275) writes [len][bytes] only: the rest of the slot keeps its old bytes."""
        self._check(fldNo)
        b = val.encode("utf-8")
        o = self.fldOffset[fldNo - 1]
        struct.pack_into(">H", self.data, o, len(b))
        self.data[o + 2:o + 2 + len(b)] = b
        return self

    def setFld(self, fldNo: int, fldBytes: bytes) -> "Tuple":
        self._check(fldNo)
        o = self.fldOffset[fldNo - 1]
        self.data[o:o + len(fldBytes)] = fldBytes
        return self

    # ---- misc ------------------------------------------------------------------------------------------
    def noOfFlds(self) -> int:
        return self.fldCnt

    def size(self) -> int:
        return self.fldOffset[self.fldCnt] - self.tuple_offset if self.fldOffset else self.tuple_length

    def getLength(self) -> int:
        return self.tuple_length

    def getTupleByteArray(self) -> bytes:
        return bytes(self.data[self.tuple_offset:self.tuple_offset + self.size()])

    returnTupleByteArray = getTupleByteArray

    def tupleCopy(self, fromTuple: "Tuple") -> None:
        b = fromTuple.getTupleByteArray()
        self.data[self.tuple_offset:self.tuple_offset + len(b)] = b
