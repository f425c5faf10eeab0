"""Mirror of bitmap.BitMapFile's in-memory surface (minijava/src/bitmap/BitMapFile.java:291-305,478) and the
java.util.BitSet operations the scan path uses.  The bits live in uint64 words in BitSet.toLongArray()
order, which is exactly what libmbcol hands out (bit p = word p/64, bit p%64 = byte p/8, bit p%8)."""
from __future__ import annotations

import numpy as np


class BitSet:
    """The subset of java.util.BitSet the reference's scan path calls."""

    def __init__(self, words=None, nbits: int = 0):
        self.words = np.zeros((nbits + 63) // 64, dtype=np.uint64) if words is None else np.array(words, dtype=np.uint64)

    def _grow(self, nwords: int) -> None:
        if nwords > self.words.size:
            self.words = np.concatenate([self.words, np.zeros(nwords - self.words.size, dtype=np.uint64)])

    def get(self, p: int) -> bool:
        w = p >> 6
        return w < self.words.size and bool((int(self.words[w]) >> (p & 63)) & 1)

    def set(self, p: int) -> None:
        self._grow((p >> 6) + 1)
        self.words[p >> 6] |= np.uint64(1 << (p & 63))

    def clear(self, p: int) -> None:
        if (p >> 6) < self.words.size:
            self.words[p >> 6] &= np.uint64(~(1 << (p & 63)) & 0xFFFFFFFFFFFFFFFF)

    def nextSetBit(self, fromIndex: int) -> int:
        w = fromIndex >> 6
        if w >= self.words.size:
            return -1
        cur = int(self.words[w]) & (~((1 << (fromIndex & 63)) - 1) & 0xFFFFFFFFFFFFFFFF)
        while True:
            if cur:
                return (w << 6) + ((cur & -cur).bit_length() - 1)
            w += 1
            if w >= self.words.size:
                return -1
            cur = int(self.words[w])

    def or_(self, other: "BitSet") -> None:
        self._grow(other.words.size)
        self.words[:other.words.size] |= other.words

    def and_(self, other: "BitSet") -> None:
        n = min(self.words.size, other.words.size)
        self.words[:n] &= other.words[:n]
        self.words[n:] = 0

    def andNot(self, other: "BitSet") -> None:
        n = min(self.words.size, other.words.size)
        self.words[:n] &= ~other.words[:n]

    def cardinality(self) -> int:
        return int(np.unpackbits(self.words.view(np.uint8)).sum())

    def length(self) -> int:
        nz = np.nonzero(self.words)[0]
        return 0 if nz.size == 0 else int(nz[-1]) * 64 + int(self.words[nz[-1]]).bit_length()

    def isEmpty(self) -> bool:
        return not self.words.any()

    def toByteArray(self) -> bytes:
        return self.words.view(np.uint8).tobytes().rstrip(b"\0")       # trailing zero bytes are dropped

    def toLongArray(self) -> np.ndarray:
        nz = np.nonzero(self.words)[0]
        return self.words[:0 if nz.size == 0 else nz[-1] + 1].copy()

    def positions(self) -> np.ndarray:
        return np.nonzero(np.unpackbits(self.words.view(np.uint8), bitorder="little"))[0].astype(np.int64)

    def __iter__(self):
        return iter(self.positions().tolist())

    def __eq__(self, other):
        return isinstance(other, BitSet) and np.array_equal(self.toLongArray(), other.toLongArray())

    def __repr__(self):                                          # BitSet.toString(): {1, 2, 5}
        return "{" + ", ".join(str(p) for p in self.positions().tolist()) + "}"


class BitMapFile:
    """One value's bitmap (or the markedDeleted bitmap).  set/clear/isSet/getBitSet as in
    bitmap/BitMapFile.java:291-305,478.  An instance obtained from Columnarfile.getBitmapIndex is a host copy of
    the device bitmap; the markedDeleted instance writes through to the device table on every change."""

    def __init__(self, bitSet: BitSet | None = None, on_change=None):
        self.bitSet = bitSet if bitSet is not None else BitSet()
        self._on_change = on_change

    def set(self, position: int) -> None:
        self.bitSet.set(position)
        if self._on_change:
            self._on_change(self.bitSet)

    def clear(self, position: int) -> None:
        self.bitSet.clear(position)
        if self._on_change:
            self._on_change(self.bitSet)

    def isSet(self, position: int) -> bool:
        return self.bitSet.get(position)

    def isClear(self, position: int) -> bool:
        return not self.bitSet.get(position)

    def getBitSet(self) -> BitSet:
        return self.bitSet
