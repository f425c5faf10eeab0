"""CPU oracle for the columnar scan path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (libmbcol.so and the minibase-columnar-database_b200 package) never does.

Parity status: PINNED against the reference's own golden transcript (phase3_output, extracted to
tests/golden/ by tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.

Pieces (paths relative to /root/reference/minijava/src):
  * scan / join         : C++ literal restatement in mbc_oracle.cpp (ctypes wrappers below)
  * synthetic tables    : the counter RNG of SURVEY.md 8d, numpy
  * DB-file writer/reader in the reference page format: diskmgr/DB.java:866-871,985-1000 (file
    directory), heap/HFPage.java:31-40,295-396 (slotted page), heap/DataPageInfo.java:19-29
    (directory records), heap/Heapfile.java:262-289,349-417 (position <-> RID),
    columnar/Columnarfile.java:60-102,257-323 (.hdr schema records), bitmap/BM.java:64-129 (.md chain)
  * bitmap index + bitmap CNF scan: columnar/Columnarfile.java:698-753,
    index/ColumnIndexScan.java:656-740, index/ColumnarIndexScan.java:130-181 (incl. its
    duplicate-constraint cache), java.util.BitSet.toByteArray sizes
  * CNF string grammar  : input/Query.java:299-323, input/BitMapQuery.java:347-420,
    input/MultiIndexQuery.java (the 4-field form with access type)
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from collections import namedtuple
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmbc_oracle.so")

ATTR_STRING, ATTR_INTEGER, ATTR_REAL, ATTR_SYMBOL = 0, 1, 2, 3
OP_EQ, OP_LT, OP_GT, OP_NE, OP_LE, OP_GE, OP_NOT, OP_NOP, OP_RANGE = range(9)
AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX = 0, 1, 2, 3
OPS = {"=": OP_EQ, "<": OP_LT, ">": OP_GT, "!=": OP_NE, ">=": OP_GE, "<=": OP_LE}       # AttrOperator.findOperator
OP_NAMES = {OP_EQ: "aopEQ", OP_LT: "aopLT", OP_GT: "aopGT", OP_NE: "aopNE", OP_LE: "aopLE", OP_GE: "aopGE",
            OP_NOT: "aopNOT", OP_NOP: "aopNOP", OP_RANGE: "opRANGE"}                     # AttrOperator.toString

# same attribute names as the product's engine.Term, so tests can hand one list to both sides
Term = namedtuple("Term", "op lhs rhs conj", defaults=(0,))


def build() -> str:
    """Compile mbc_oracle.cpp (g++) if needed; returns the library path."""
    src = os.path.join(_HERE, "mbc_oracle.cpp")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


# ---------------------------------------------------------------------------------------------
# ctypes view of mbc_oracle.cpp
# ---------------------------------------------------------------------------------------------
class _OCol(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_int32), ("data", C.c_void_p)]


class _OOperand(C.Structure):
    _fields_ = [("kind", C.c_int32), ("type", C.c_int32), ("col", C.c_int32), ("lit_i", C.c_int32),
                ("lit_f", C.c_float), ("lit_slen", C.c_int32), ("lit_s", C.POINTER(C.c_uint8))]


class _OTerm(C.Structure):
    _fields_ = [("op", C.c_int32), ("conj_id", C.c_int32), ("lhs", _OOperand), ("rhs", _OOperand)]


class _OAgg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32)]


class _OAggOut(C.Structure):
    _fields_ = [("i", C.c_int64), ("f", C.c_double), ("valid", C.c_int32), ("pad", C.c_int32)]


class _OProj(C.Structure):
    _fields_ = [("rel", C.c_int32), ("col", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_scan.restype = C.c_int64
        _lib.orc_bitmap_join.restype = C.c_int64
        _lib.orc_tuple_len.restype = C.c_int32
        _lib.orc_max_threads.restype = C.c_int32
    return _lib


def max_threads() -> int:
    return int(_load().orc_max_threads())


def _fill_operand(o, spec, keep):
    kind, val = spec
    o.kind = o.type = o.col = o.lit_i = o.lit_slen = 0
    o.lit_f = 0.0
    if kind == "col":
        o.kind, o.type, o.col = 1, ATTR_SYMBOL, int(val)
    elif kind == "icol":
        o.kind, o.type, o.col = 2, ATTR_SYMBOL, int(val)
    elif kind == "int":
        o.kind, o.type, o.lit_i = 0, ATTR_INTEGER, int(val)
    elif kind == "real":
        o.kind, o.type, o.lit_f = 0, ATTR_REAL, float(np.float32(val))
    elif kind == "str":
        b = val.encode("utf-8") if isinstance(val, str) else bytes(val)
        buf = (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b if b else b"\0")
        keep.append(buf)
        o.kind, o.type, o.lit_slen, o.lit_s = 0, ATTR_STRING, len(b), C.cast(buf, C.POINTER(C.c_uint8))
    else:
        raise ValueError(kind)


def _pack_terms(terms):
    terms = sorted(terms, key=lambda t: t.conj)
    arr = (_OTerm * max(len(terms), 1))()
    keep = []
    for i, t in enumerate(terms):
        arr[i].op, arr[i].conj_id = int(t.op), int(t.conj)
        _fill_operand(arr[i].lhs, t.lhs, keep)
        _fill_operand(arr[i].rhs, t.rhs, keep)
    return arr, len(terms), keep


def _pack_cols(coldescs, columns):
    arrs = []
    cols = (_OCol * len(coldescs))()
    for c, ((t, w), a) in enumerate(zip(coldescs, columns)):
        if t == ATTR_INTEGER:
            a = np.ascontiguousarray(a, dtype=np.int32)
        elif t == ATTR_REAL:
            a = np.ascontiguousarray(a, dtype=np.float32)
        else:
            a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
        arrs.append(a)
        cols[c].type, cols[c].width, cols[c].data = int(t), int(w), a.ctypes.data
    return cols, arrs


def nrows_of(coldescs, columns) -> int:
    t, w = coldescs[0]
    a = np.asarray(columns[0])
    return a.size // w if t == ATTR_STRING else a.size


def scan(coldescs, columns, terms: Sequence = (), proj: Sequence[int] = (), aggs: Sequence[tuple] = (),
         deleted_words=None, want_positions=True, want_tuples=True, stale_padding=False, nthreads=1) -> dict:
    """ColumnarFileScan restated: positions (ascending), projected tuples (reference bytes), aggregates."""
    lib = _load()
    n = nrows_of(coldescs, columns)
    cols, keep_cols = _pack_cols(coldescs, columns)
    tarr, nt, keep = _pack_terms(terms)
    parr = (C.c_int32 * max(len(proj), 1))(*proj)
    aarr = (_OAgg * max(len(aggs), 1))(*[_OAgg(int(k), int(c)) for k, c in aggs])
    aout = (_OAggOut * max(len(aggs), 1))()
    tlen = int(lib.orc_tuple_len(len(coldescs), cols, parr, len(proj)))
    pos = np.empty(n if want_positions else 0, dtype=np.int64)
    tup = np.empty((n if (want_tuples and proj) else 0) * tlen, dtype=np.uint8)
    dw = None
    if deleted_words is not None:
        dw = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
        src = np.asarray(deleted_words, dtype=np.uint64)
        dw[:min(src.size, dw.size)] = src[:dw.size]
    cnt = lib.orc_scan(C.c_int32(len(coldescs)), cols, C.c_int64(n), C.c_void_p(dw.ctypes.data if dw is not None else None),
                       tarr, C.c_int32(nt), parr, C.c_int32(len(proj)),
                       C.c_void_p(pos.ctypes.data if want_positions else None),
                       C.c_void_p(tup.ctypes.data if (want_tuples and proj) else None),
                       C.c_int32(1 if stale_padding else 0), aarr, C.c_int32(len(aggs)), aout, C.c_int32(nthreads))
    del keep, keep_cols
    return {
        "count": int(cnt),
        "positions": pos[:cnt] if want_positions else None,
        "tuples": tup[:cnt * tlen].reshape(cnt, tlen) if (want_tuples and proj) else None,
        "tuple_len": tlen,
        "aggs": [(int(aout[a].i), float(aout[a].f), bool(aout[a].valid)) for a in range(len(aggs))],
    }


def bitmap_join(ocoldescs, ocolumns, icoldescs, icolumns, join_terms, proj, aggs=(), outer_sel=None, inner_sel=None,
                outer_deleted=None, inner_deleted=None) -> dict:
    """BitMapQuery.executeJoin restated.  proj = [(rel, col)], rel 1 = outer, 2 = inner."""
    lib = _load()
    no, ni = nrows_of(ocoldescs, ocolumns), nrows_of(icoldescs, icolumns)
    oc, k1 = _pack_cols(ocoldescs, ocolumns)
    ic, k2 = _pack_cols(icoldescs, icolumns)
    tarr, nt, keep = _pack_terms(join_terms)
    parr = (_OProj * max(len(proj), 1))(*[_OProj(int(r), int(c)) for r, c in proj])
    aarr = (_OAgg * max(len(aggs), 1))(*[_OAgg(int(k), int(c)) for k, c in aggs])
    aout = (_OAggOut * max(len(aggs), 1))()

    def words(w, n):
        if w is None:
            return None
        out = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
        src = np.asarray(w, dtype=np.uint64)
        out[:min(src.size, out.size)] = src[:out.size]
        return out

    osel, isel, odel, idel = words(outer_sel, no), words(inner_sel, ni), words(outer_deleted, no), words(inner_deleted, ni)
    ptr = lambda a: C.c_void_p(a.ctypes.data if a is not None else None)
    descs = [(ocoldescs if r == 1 else icoldescs)[c] for r, c in proj]
    tlen = (len(descs) + 2) * 2 + sum(w + 2 if t == ATTR_STRING else 4 for t, w in descs)

    def run(cap, op, ip, tp):
        return lib.orc_bitmap_join(C.c_int32(len(ocoldescs)), oc, C.c_int64(no), ptr(osel), ptr(odel),
                                   C.c_int32(len(icoldescs)), ic, C.c_int64(ni), ptr(isel), ptr(idel),
                                   tarr, C.c_int32(nt), parr, C.c_int32(len(proj)), C.c_int64(cap), ptr(op), ptr(ip), ptr(tp),
                                   aarr, C.c_int32(len(aggs)), aout)

    cnt = run(0, None, None, None)
    opos = np.empty(cnt, dtype=np.int64)
    ipos = np.empty(cnt, dtype=np.int64)
    tup = np.empty(cnt * tlen, dtype=np.uint8)
    run(cnt, opos, ipos, tup)
    del keep, k1, k2
    return {"count": int(cnt), "outer_positions": opos, "inner_positions": ipos, "tuples": tup.reshape(cnt, tlen),
            "tuple_len": tlen, "aggs": [(int(aout[a].i), float(aout[a].f), bool(aout[a].valid)) for a in range(len(aggs))]}


# ---------------------------------------------------------------------------------------------
# synthetic tables -- must match csrc/mbc_synth.cu bit for bit
# ---------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _draw(seed: int, col: int, pos: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = np.uint64((seed ^ (((col + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF)
        return splitmix64(key + pos.astype(np.uint64))


def synth_int(seed: int, col: int, nrows: int, domain: int, position_base: int = 0) -> np.ndarray:
    pos = np.arange(position_base, position_base + nrows, dtype=np.uint64)
    return (_draw(seed, col, pos) % np.uint64(domain)).astype(np.uint32).view(np.int32)


def synth_real(seed: int, col: int, nrows: int, position_base: int = 0) -> np.ndarray:
    pos = np.arange(position_base, position_base + nrows, dtype=np.uint64)
    d = (_draw(seed, col, pos) >> np.uint64(40)).astype(np.float64) * (1000.0 / 16777216.0)
    return d.astype(np.float32)


def synth_perm(nrows: int, domain: int, position_base: int = 0) -> np.ndarray:
    pos = np.arange(position_base, position_base + nrows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return ((pos * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(domain)).astype(np.uint32).view(np.int32)


def synth_str(seed: int, col: int, nrows: int, width: int, position_base: int = 0) -> np.ndarray:
    pos = np.arange(position_base, position_base + nrows, dtype=np.uint64)
    x = _draw(seed, col, pos)
    out = np.zeros((nrows, width), dtype=np.uint8)
    with np.errstate(over="ignore"):
        for blk in range((width + 7) // 8):
            y = splitmix64(x ^ np.uint64(((blk + 1) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF))
            for k in range(blk * 8, min(width, blk * 8 + 8)):
                b = (y >> np.uint64(8 * (k & 7))) & np.uint64(0xFF)
                out[:, k] = (np.uint64(0x21) + b % np.uint64(94)).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------
# strings <-> fixed-width zero padded byte rows
# ---------------------------------------------------------------------------------------------
def nlj_outer_block_rows(outer_coldescs, outer_cols_in_tuple: Sequence[int], amt_of_mem: int) -> int:
    """Outer tuples per block of iterator/ColumnarNestedLoopJoins.java:118-121: (amt_of_mem - 1) buffer pages of
    MINIBASE_PAGESIZE / outerItr.getTupleSize() tuples, the tuple holding the columns input/NljQuery.java collects for the
    outer side (targets, join columns, constraint columns of the conjuncts the scan does not evaluate; a TreeSet)."""
    cols = sorted(set(outer_cols_in_tuple))
    size = (len(cols) + 2) * 2 + sum(outer_coldescs[c][1] + 2 if outer_coldescs[c][0] == ATTR_STRING else 4 for c in cols)
    return (amt_of_mem - 1) * (1024 // size)


def nlj_order(outer_qualifying_positions, pair_outer, pair_inner, block_rows: int) -> np.ndarray:
    """Emission order of ColumnarNestedLoopJoins.get_next (:157-207) as a permutation of the pair list: the qualifying
    outer rows are taken block_rows at a time (scan order); for every block the whole inner side is walked in order, and
    for every inner row the block's outer rows in order.  Sort key = (outer block, inner position, outer position)."""
    oq = np.asarray(outer_qualifying_positions, dtype=np.int64)
    po, pi = np.asarray(pair_outer, dtype=np.int64), np.asarray(pair_inner, dtype=np.int64)
    rank = np.searchsorted(oq, po)
    return np.lexsort((po, pi, rank // max(int(block_rows), 1)))


def external_sort_order(coldescs, columns, key_cols: Sequence[int], descending: bool, sort_buffers: int) -> list:
    """input/ColumnarSort.java:206-352 restated on row ids: the ORDER (ties included) in which the reference's external
    merge sort emits the rows.  (key..., position) records of `rec` bytes go P = 1004 // (rec + 4) to a heap page
    (:206-210, heap/HFPage.java); B = sort_buffers - 1 input buffers; ceil(log(pages) / log(B)) passes (:214).  Pass 0
    takes the pages B at a time, reads them round-robin (first record of every page, second record of every page, ...,
    :293-309) and sorts that list stably (Collections.sort, :310); every later pass merges B runs of B^i pages, always
    taking the smallest head and, among equal heads, the one of the lowest run (:318-346,664-680: a stable sort of the
    (run, head) entries).  Every pass appends to a fresh heap file, so a run starts on a page boundary."""
    import math
    n = nrows_of(coldescs, columns)
    ranks = []
    for c in key_cols:
        t, w = coldescs[c]
        col = np.asarray(columns[c])
        keys = np.ascontiguousarray(col.reshape(n, -1)[:, :w]).view(f"S{w}").reshape(n) if t == ATTR_STRING else col.reshape(n)
        _, inv = np.unique(keys, return_inverse=True)
        ranks.append(-inv if descending else inv)
    key = [tuple(int(r[i]) for r in ranks) for i in range(n)]
    rec = sum(coldescs[c][1] + 2 if coldescs[c][0] == ATTR_STRING else 4 for c in key_cols) + 4
    per_page = 1004 // (rec + 4)
    nbuf = sort_buffers - 1
    npages = (n + per_page - 1) // per_page
    passes = int(math.ceil(math.log(npages) / math.log(nbuf))) if npages > 1 else 0
    cur = list(range(n))
    for i in range(passes):
        pages = [cur[s:s + per_page] for s in range(0, n, per_page)]
        run = nbuf ** i
        runs = [pages[s:s + run] for s in range(0, npages, run)]
        out = []
        for k in range(0, len(runs), nbuf):
            group = runs[k:k + nbuf]
            if i == 0:
                heads = [r[0] for r in group]
                lst = [h[s] for s in range(max(len(h) for h in heads)) for h in heads if s < len(h)]
                lst.sort(key=lambda r: key[r])
                out += lst
            else:
                seqs = [[r for p in rn for r in p] for rn in group]
                ptr = [0] * len(seqs)
                while True:
                    best = -1
                    for j, sq in enumerate(seqs):
                        if ptr[j] < len(sq) and (best < 0 or key[sq[ptr[j]]] < key[seqs[best][ptr[best]]]):
                            best = j
                    if best < 0:
                        break
                    out.append(seqs[best][ptr[best]])
                    ptr[best] += 1
        cur = out
    return cur


def sort(coldescs, columns, key_cols: Sequence[int], descending: bool = False, deleted_positions=()) -> np.ndarray:
    """input/ColumnarSort.java:163-205 comparator, restated: the live positions ordered by the key columns (first key most
    significant; ints numerically, strings by String.compareTo = byte order of the zero-padded BMP text), all keys
    ascending or all descending.  Ties keep ascending position (the Java's external merge sort leaves tie order
    unspecified: its golden output is compared as key sequence + row multiset)."""
    n = nrows_of(coldescs, columns)
    if n == 0:
        return np.empty(0, dtype=np.int64)
    live = np.ones(n, dtype=bool)
    live[list(deleted_positions)] = False
    ranks = []
    for c in key_cols:
        t, w = coldescs[c]
        col = np.asarray(columns[c])
        if t == 0:
            keys = np.ascontiguousarray(col.reshape(n, -1)[:, :w]).view(f"S{w}").reshape(n)     # bytes compare, NUL padded
        else:
            keys = col.reshape(n)
        _, inv = np.unique(keys, return_inverse=True)                                            # dense rank per column
        ranks.append(-inv if descending else inv)
    order = np.lexsort(tuple(reversed(ranks))) if ranks else np.arange(n)                        # stable: ties by position
    return order[live[order]].astype(np.int64)


def pack_strings(values: Sequence[str], width: int) -> np.ndarray:
    out = np.zeros((len(values), width), dtype=np.uint8)
    for i, s in enumerate(values):
        b = s.encode("utf-8")
        if len(b) > width:
            raise ValueError(f"{s!r} longer than char({width})")
        out[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    return out


def unpack_strings(rows: np.ndarray) -> list[str]:
    return [bytes(r).rstrip(b"\0").decode("utf-8") for r in np.asarray(rows, dtype=np.uint8)]


def load_tsv(path: str):
    """BatchInsert.java:60-104: header `name:int` / `name:char(N)`, tab separated rows."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    names, coldescs = [], []
    for h in lines[0].split("\t"):
        name, typ = h.split(":")
        names.append(name)
        coldescs.append((ATTR_INTEGER, 4) if typ == "int" else (ATTR_STRING, int(typ[typ.index("(") + 1:typ.index(")")])))
    raw = [ln.split("\t") for ln in lines[1:]]
    columns = []
    for c, (t, w) in enumerate(coldescs):
        vals = [r[c] for r in raw]
        columns.append(np.array([int(v) for v in vals], dtype=np.int32) if t == ATTR_INTEGER else pack_strings(vals, w))
    return names, coldescs, columns


# ---------------------------------------------------------------------------------------------
# reference Tuple bytes (heap/Tuple.java:369-440)
# ---------------------------------------------------------------------------------------------
def tuple_layout(descs):
    n = len(descs)
    offs = [(n + 2) * 2]
    for t, w in descs:
        offs.append(offs[-1] + (w + 2 if t == ATTR_STRING else 4))
    return offs


def decode_tuple(buf: bytes, descs) -> list:
    """Read the fields the way Tuple.getIntFld/getFloFld/getStrFld do (through fldOffset[])."""
    n = struct.unpack(">h", buf[0:2])[0]
    offs = [struct.unpack(">h", buf[2 + 2 * i:4 + 2 * i])[0] for i in range(n + 1)]
    out = []
    for i, (t, w) in enumerate(descs):
        o = offs[i]
        if t == ATTR_INTEGER:
            out.append(struct.unpack(">i", buf[o:o + 4])[0])
        elif t == ATTR_REAL:
            out.append(struct.unpack(">f", buf[o:o + 4])[0])
        else:
            ln = struct.unpack(">H", buf[o:o + 2])[0]
            out.append(bytes(buf[o + 2:o + 2 + ln]).decode("utf-8"))
    return out


# ---------------------------------------------------------------------------------------------
# java.util.BitSet views over uint64 words
# ---------------------------------------------------------------------------------------------
def bits_from_positions(positions, nbits: int) -> np.ndarray:
    w = np.zeros((nbits + 63) // 64, dtype=np.uint64)
    p = np.asarray(positions, dtype=np.int64)
    np.bitwise_or.at(w, p >> 6, np.uint64(1) << (p & 63).astype(np.uint64))
    return w


def positions_from_bits(words: np.ndarray, nbits: int | None = None) -> np.ndarray:
    b = np.unpackbits(np.ascontiguousarray(words, dtype=np.uint64).view(np.uint8), bitorder="little")
    if nbits is not None:
        b = b[:nbits]
    return np.nonzero(b)[0].astype(np.int64)


def bitset_bytearray_len(words: np.ndarray) -> int:
    """len(BitSet.toByteArray()): floor(highest set bit / 8) + 1, 0 when empty."""
    p = positions_from_bits(words)
    return 0 if p.size == 0 else int(p[-1]) // 8 + 1


def mask_to_words(mask: np.ndarray) -> np.ndarray:
    n = mask.size
    padded = np.zeros(((n + 63) // 64) * 64, dtype=np.uint8)
    padded[:n] = mask.astype(np.uint8)
    return np.packbits(padded, bitorder="little").view(np.uint64).copy()


# ---------------------------------------------------------------------------------------------
# bitmap index + bitmap CNF scan
# ---------------------------------------------------------------------------------------------
def column_values(desc, column) -> list:
    t, w = desc
    if t == ATTR_STRING:
        return unpack_strings(np.asarray(column, dtype=np.uint8).reshape(-1, w))
    return [int(v) for v in np.asarray(column)]


def bitmap_build(desc, column, deleted_words=None) -> dict:
    """Columnarfile.createBitMapIndex (:698-753): value -> bitset words; rows deleted at build time are skipped."""
    t, w = desc
    n = np.asarray(column).size // (w if t == ATTR_STRING else 1)
    alive = np.ones(n, dtype=bool)
    if deleted_words is not None:
        alive[positions_from_bits(deleted_words, n)] = False
    out = {}
    if t == ATTR_STRING:
        rows = np.asarray(column, dtype=np.uint8).reshape(n, w)
        uniq, inv = np.unique(rows, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        for k in range(uniq.shape[0]):
            m = (inv == k) & alive
            if m.any():
                out[bytes(uniq[k]).rstrip(b"\0").decode("utf-8")] = mask_to_words(m)
    else:
        col = np.asarray(column)
        for v in np.unique(col[alive]):
            out[int(v)] = mask_to_words((col == v) & alive)
    return out


def _java_compare(a, b) -> int:
    """Integer compare, or String.compareTo (UTF-16 code units)."""
    if isinstance(a, str):
        ua, ub = a.encode("utf-16-be"), b.encode("utf-16-be")
        ka = [int.from_bytes(ua[i:i + 2], "big") for i in range(0, len(ua), 2)]
        kb = [int.from_bytes(ub[i:i + 2], "big") for i in range(0, len(ub), 2)]
        for x, y in zip(ka, kb):
            if x != y:
                return x - y
        return len(ka) - len(kb)
    return (a > b) - (a < b)


def bitmap_term_bits(index: dict, op: int, literal, nbits: int) -> np.ndarray:
    """ColumnIndexScan.getBitSet (:656-740): OR of the bitmaps of the indexed values selected by `col op literal`."""
    out = np.zeros((nbits + 63) // 64, dtype=np.uint64)
    chosen = []
    if op in (OP_EQ, OP_LE, OP_GE):
        chosen.append(literal)
    if op in (OP_LT, OP_LE):
        chosen += [v for v in index if _java_compare(literal, v) > 0]
    if op in (OP_GT, OP_GE):
        chosen += [v for v in index if _java_compare(literal, v) < 0]
    if op == OP_NE:
        chosen += [v for v in index if _java_compare(literal, v) != 0]
    for v in chosen:                       # getBitmapIndex of a value never indexed = empty BitMapFile
        if v in index:
            out |= index[v]
    return out


def bitmap_cnf(indexes: dict, colnames: Sequence[str], conjuncts, nbits: int, deleted_words=None,
               emulate_duplicate_cache: bool = True) -> np.ndarray:
    """ColumnarIndexScan constructor (:130-181).

    conjuncts: list of lists of (col, op, literal[, access]) ; access in {"BM","BT"} defaults to "BM".
    indexes[col] = bitmap_build(...) result.  Every term is evaluated through the bitmap semantics
    (a B-tree term selects the same positions).  With emulate_duplicate_cache the method's
    duplicateConstraints HashMap is reproduced: a term whose text occurs more than once in the
    query string is cached the first time it is evaluated -- and what is cached is the *conjunct's
    accumulating* BitSet object (``positions``), not the term's own bits (:147-172).
    """
    idx_name = {"BM": "Bitmap", "BT": "B_Index"}

    def text(term):
        col, op, lit = term[0], term[1], term[2]
        acc = term[3] if len(term) > 3 else "BM"
        return f"{colnames[col]}{OP_NAMES[op]}{lit}{idx_name[acc]}"

    query = "^".join("|".join(text(t) for t in conj) for conj in conjuncts)      # buildInputQueryString (:330-350)
    nwords = (nbits + 63) // 64
    deleted = np.zeros(nwords, dtype=np.uint64)
    if deleted_words is not None:
        src = np.asarray(deleted_words, dtype=np.uint64)
        deleted[:min(nwords, src.size)] = src[:nwords]
    cache: dict[str, np.ndarray] = {}
    output = None
    for i, conj in enumerate(conjuncts):
        positions = np.zeros(nwords, dtype=np.uint64)             # one BitSet object per conjunct
        for term in conj:
            key = text(term)
            dup = emulate_duplicate_cache and (key in cache or _count_occurrences(query, key) > 1)
            if not dup or key not in cache:
                bits = bitmap_term_bits(indexes[term[0]], term[1], term[2], nbits) & ~deleted   # get_bm_next_tid skips deleted
                positions |= bits                                  # in place: cached aliases see it
                if dup:
                    cache[key] = positions                         # the conjunct's BitSet, by reference
            else:
                positions |= cache[key]
        if i == 0:
            output = positions                                     # outputPositions = positions (same object)
        else:
            output &= positions
    return output


def _count_occurrences(hay: str, needle: str) -> int:            # checkDuplicateConstraint (:352-369)
    count, i = 0, 0
    while True:
        i = hay.find(needle, i)
        if i < 0:
            return count
        count += 1
        i += len(needle)


# ---------------------------------------------------------------------------------------------
# CNF strings
# ---------------------------------------------------------------------------------------------
def parse_cnf(s: str, colnames: Sequence[str], coldescs) -> list:
    """`{(A,=,x)|(B,<,y)}^{(C,!=,6)}` -> [[(col, op, literal[, access])...]...]; literal typed by the column
    (Query.java:311-321)."""
    out = []
    for conj in s.strip().split("^"):
        conj = conj.strip()
        if not (conj.startswith("{") and conj.endswith("}")):
            raise ValueError("Invalid query format")
        terms = []
        for dis in conj[1:-1].split("|"):
            dis = dis.strip()
            if not (dis.startswith("(") and dis.endswith(")")):
                raise ValueError("Invalid query format")
            parts = [p.strip() for p in dis[1:-1].strip().split(",")]
            if len(parts) not in (3, 4):
                raise ValueError("Invalid VALUECONSTRAINT elements")
            col = list(colnames).index(parts[0])
            lit = int(parts[2]) if coldescs[col][0] == ATTR_INTEGER else parts[2]
            terms.append((col, OPS[parts[1]], lit) + ((parts[3],) if len(parts) == 4 else ()))
        out.append(terms)
    return out


def cnf_to_terms(conjuncts, coldescs) -> list:
    """Parsed CNF -> Term list for scan() (column op literal, literal typed by the column)."""
    terms = []
    for ci, conj in enumerate(conjuncts):
        for t in conj:
            col, op, lit = t[0], t[1], t[2]
            kind = {ATTR_INTEGER: "int", ATTR_REAL: "real", ATTR_STRING: "str"}[coldescs[col][0]]
            terms.append(Term(op, ("col", col), (kind, lit), ci))
    return terms


# ---------------------------------------------------------------------------------------------
# reference DB file images
# ---------------------------------------------------------------------------------------------
PAGE = 1024
DPFIXED = 20
INVALID_PAGE = -1
MAX_NAME = 50
FILE_ENTRY = 4 + MAX_NAME + 2
DIR_RECS = (PAGE - DPFIXED) // (4 + 8)          # DataPageInfo records per directory page = 83


def _utf(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack(">H", len(b)) + b


class _HFPage:
    """heap/HFPage.java: 20-byte header, slot directory growing up, records packed down from 1024."""

    def __init__(self, pid: int):
        self.b = bytearray(PAGE)
        self.pid = pid
        struct.pack_into(">hhhhiii", self.b, 0, 0, PAGE, PAGE - DPFIXED, 0, INVALID_PAGE, INVALID_PAGE, pid)

    @property
    def free(self):
        return struct.unpack_from(">h", self.b, 4)[0]

    @property
    def slot_cnt(self):
        return struct.unpack_from(">h", self.b, 0)[0]

    def set_next(self, pid):
        struct.pack_into(">i", self.b, 12, pid)

    def set_prev(self, pid):
        struct.pack_into(">i", self.b, 8, pid)

    def insert(self, rec: bytes):
        need = len(rec) + 4
        if need > self.free:
            return None
        cnt, used, free = struct.unpack_from(">hhh", self.b, 0)
        used -= len(rec)
        struct.pack_into(">hhh", self.b, 0, cnt + 1, used, free - need)
        struct.pack_into(">hh", self.b, DPFIXED + 4 * cnt, len(rec), used)
        self.b[used:used + len(rec)] = rec
        return cnt

    def update(self, slot: int, rec: bytes):
        ln, off = struct.unpack_from(">hh", self.b, DPFIXED + 4 * slot)
        assert ln == len(rec)
        self.b[off:off + ln] = rec


class DBWriter:
    """Builds a DB file image the way the reference lays it out (diskmgr/DB.java): page 0 = first
    directory page, then the space map, then file pages in allocation order."""

    def __init__(self, num_pages: int = 1 << 20):
        self.num_pages = num_pages
        self.pages: dict[int, bytearray] = {}
        nmap = (num_pages + PAGE * 8 - 1) // (PAGE * 8)
        self.next_free = 1 + nmap
        p0 = bytearray(PAGE)
        n0 = (PAGE - 20) // FILE_ENTRY
        struct.pack_into(">ii", p0, 0, INVALID_PAGE, n0)
        for i in range(n0):
            struct.pack_into(">i", p0, 8 + i * FILE_ENTRY, INVALID_PAGE)
        struct.pack_into(">i", p0, PAGE - 4, num_pages)
        self.pages[0] = p0
        self.dir_pages = [0]

    def alloc(self) -> int:
        pid = self.next_free
        self.next_free += 1
        return pid

    def add_file_entry(self, name: str, first_pid: int):
        if len(name.encode()) > MAX_NAME:
            raise ValueError("file name too long")
        for dp in self.dir_pages:
            b = self.pages[dp]
            n = struct.unpack_from(">i", b, 4)[0]
            for i in range(n):
                if struct.unpack_from(">i", b, 8 + i * FILE_ENTRY)[0] == INVALID_PAGE:
                    struct.pack_into(">i", b, 8 + i * FILE_ENTRY, first_pid)
                    u = _utf(name)
                    b[8 + i * FILE_ENTRY + 4:8 + i * FILE_ENTRY + 4 + len(u)] = u
                    return
        # directory full: chain a new directory page (DB.java add_file_entry)
        pid = self.alloc()
        b = bytearray(PAGE)
        n = (PAGE - 16) // FILE_ENTRY
        struct.pack_into(">ii", b, 0, INVALID_PAGE, n)
        for i in range(n):
            struct.pack_into(">i", b, 8 + i * FILE_ENTRY, INVALID_PAGE)
        struct.pack_into(">i", self.pages[self.dir_pages[-1]], 0, pid)
        self.pages[pid] = b
        self.dir_pages.append(pid)
        self.add_file_entry(name, first_pid)

    def tobytes(self) -> bytes:
        top = self.next_free
        out = bytearray(top * PAGE)
        for pid, b in self.pages.items():
            out[pid * PAGE:(pid + 1) * PAGE] = b
        # the space map (pages 1..): bit p of the map = page p allocated, least significant bit first (DB.java:739-822);
        # openDB marks page 0 and the map pages (:104-106), allocate_page every page handed out since
        for pid in range(top):
            at = PAGE * (1 + pid // (PAGE * 8)) + (pid % (PAGE * 8)) // 8
            out[at] |= 1 << (pid % 8)
        return bytes(out)


class HeapfileWriter:
    """heap/Heapfile.java insertRecord for append-only use: directory pages of DataPageInfo records,
    each pointing at one data page."""

    def __init__(self, db: DBWriter, name: str):
        self.db = db
        first = _HFPage(db.alloc())
        db.pages[first.pid] = first.b
        db.add_file_entry(name, first.pid)
        self.dirs = [first]
        self.cur_data = None
        self.cur_info = None          # (dir page, slot)
        self.recct = 0

    def insert(self, rec: bytes):
        if self.cur_data is None or self.cur_data.free < len(rec) + 4:
            data = _HFPage(self.db.alloc())
            self.db.pages[data.pid] = data.b
            d = self.dirs[-1]
            info = struct.pack(">hhi", data.free - 4, 0, data.pid)      # DataPageInfo.availspace = HFPage.available_space() (Heapfile.java:89)
            slot = d.insert(info)
            if slot is None:
                nd = _HFPage(self.db.alloc())
                self.db.pages[nd.pid] = nd.b
                d.set_next(nd.pid)
                nd.set_prev(d.pid)
                self.dirs.append(nd)
                d = nd
                slot = d.insert(info)
            self.cur_data, self.cur_info, self.recct = data, (d, slot), 0
        slot = self.cur_data.insert(rec)
        self.recct += 1
        d, s = self.cur_info
        d.update(s, struct.pack(">hhi", self.cur_data.free - 4, self.recct, self.cur_data.pid))
        return self.cur_data.pid, slot


def write_columnar_file(db: DBWriter, name: str, colnames, coldescs, columns, deleted_positions=()):
    """Columnarfile(name, n, types, sizes, names) + insertTuple per row (Columnarfile.java:43-140,405-488):
    <name>.hdr, <name>.<i> per column, <name>.md (deleted bitmap), <name>.dtid; rows are inserted
    row by row so the data pages of the columns interleave in allocation order like the original."""
    n = len(coldescs)
    nrows = nrows_of(coldescs, columns)
    hdr = HeapfileWriter(db, name + ".hdr")
    hdr.insert(struct.pack(">i", n))
    hdr.insert(b"".join(struct.pack(">i", t) for t, _ in coldescs))
    hdr.insert(b"".join(struct.pack(">i", w) for _, w in coldescs))
    names = bytearray(n * 17)
    for i, nm in enumerate(colnames):
        u = _utf(nm)
        names[i * 17:i * 17 + len(u)] = u
    hdr.insert(bytes(names))
    hdr.insert(bytes(n))          # bTreeExist
    hdr.insert(bytes(n))          # bitmapExist
    heaps = [HeapfileWriter(db, f"{name}.{i}") for i in range(n)]
    # <name>.md : BitMapFile(name, create=true): header page with one 1000-byte record (BitMapFile.java:70-79)
    md = _HFPage(db.alloc())
    db.pages[md.pid] = md.b
    db.add_file_entry(name + ".md", md.pid)
    dele = bits_from_positions(list(deleted_positions), max(nrows, 1)) if len(deleted_positions) else np.zeros(1, np.uint64)
    dbytes = dele.view(np.uint8).tobytes().rstrip(b"\0")          # BitSet.toByteArray drops trailing zero bytes
    chunks = [dbytes[i:i + 1000] for i in range(0, len(dbytes), 1000)] or [b""]
    md.insert(chunks[0].ljust(1000, b"\0"))
    prev = md
    for ch in chunks[1:]:                                           # BM.insertBitSet chains further pages
        pg = _HFPage(db.alloc())
        db.pages[pg.pid] = pg.b
        pg.insert(ch.ljust(1000, b"\0"))
        prev.set_next(pg.pid)
        pg.set_prev(prev.pid)
        prev = pg
    HeapfileWriter(db, name + ".dtid")
    arrs = []
    for (t, w), col in zip(coldescs, columns):
        if t == ATTR_STRING:
            arrs.append(np.asarray(col, dtype=np.uint8).reshape(nrows, w))
        elif t == ATTR_INTEGER:
            arrs.append(np.asarray(col, dtype=np.int32).astype(">i4").tobytes())
        else:
            arrs.append(np.asarray(col, dtype=np.float32).astype(">f4").tobytes())
    for r in range(nrows):
        for c, (t, w) in enumerate(coldescs):
            if t == ATTR_STRING:
                raw = bytes(arrs[c][r]).rstrip(b"\0")
                rec = (struct.pack(">H", len(raw)) + raw).ljust(w + 2, b"\0")
            else:
                rec = arrs[c][4 * r:4 * r + 4]
            heaps[c].insert(rec)


def _file_entries(db: bytes) -> dict:
    out, pid = {}, 0
    while pid != INVALID_PAGE:
        base = pid * PAGE
        nxt, n = struct.unpack_from(">ii", db, base)
        for i in range(n):
            o = base + 8 + i * FILE_ENTRY
            first = struct.unpack_from(">i", db, o)[0]
            if first != INVALID_PAGE:
                ln = struct.unpack_from(">H", db, o + 4)[0]
                out[bytes(db[o + 6:o + 6 + ln]).decode("utf-8")] = first
        pid = nxt
    return out


def _heap_records(db: bytes, first_dir: int):
    """Yield (position_page_index, slot, record bytes) in Scan order; page index as Heapfile.loadPositionBuffer
    computes it (dirIdx * 83 + dirSlot)."""
    dpid, didx = first_dir, 0
    while dpid != INVALID_PAGE:
        base = dpid * PAGE
        cnt = struct.unpack_from(">h", db, base)[0]
        nxt = struct.unpack_from(">i", db, base + 12)[0]
        for s in range(cnt):
            ln, off = struct.unpack_from(">hh", db, base + DPFIXED + 4 * s)
            if ln < 0:
                continue
            data_pid = struct.unpack_from(">i", db, base + off + 4)[0]
            pbase = data_pid * PAGE
            pcnt = struct.unpack_from(">h", db, pbase)[0]
            for ps in range(pcnt):
                rl, ro = struct.unpack_from(">hh", db, pbase + DPFIXED + 4 * ps)
                if rl < 0:
                    continue
                yield didx * DIR_RECS + s, ps, bytes(db[pbase + ro:pbase + ro + rl])
        dpid = nxt
        didx += 1


def read_columnar_file(db: bytes, name: str) -> dict:
    """Columnarfile(String) open (:239-359) + a full TupleScan decode of every column, positions by
    Heapfile.findPosition arithmetic.  Returns colnames, coldescs, columns, deleted words."""
    files = _file_entries(db)
    if name + ".hdr" not in files:
        raise KeyError("Columnar File does not exist.")
    recs = [r for _, _, r in _heap_records(db, files[name + ".hdr"])]
    n = struct.unpack(">i", recs[0][:4])[0]
    types = [struct.unpack_from(">i", recs[1], 4 * i)[0] for i in range(n)]
    sizes = [struct.unpack_from(">i", recs[2], 4 * i)[0] for i in range(n)]
    colnames = []
    for i in range(n):
        ln = struct.unpack_from(">H", recs[3], 17 * i)[0]
        colnames.append(recs[3][17 * i + 2:17 * i + 2 + ln].decode("utf-8"))
    coldescs = list(zip(types, sizes))
    columns = []
    for c, (t, w) in enumerate(coldescs):
        rec_size = w + 2 if t == ATTR_STRING else w
        per_page = (PAGE - DPFIXED) // (4 + rec_size)
        items = [(pg * per_page + s, r) for pg, s, r in _heap_records(db, files[f"{name}.{c}"])]
        nrows = (max(p for p, _ in items) + 1) if items else 0
        if t == ATTR_STRING:
            col = np.zeros((nrows, w), dtype=np.uint8)
            for p, r in items:
                ln = struct.unpack(">H", r[:2])[0]
                col[p, :ln] = np.frombuffer(r[2:2 + ln], dtype=np.uint8)
        else:
            col = np.zeros(nrows, dtype=np.int32 if t == ATTR_INTEGER else np.float32)
            for p, r in items:
                col[p] = struct.unpack(">i" if t == ATTR_INTEGER else ">f", r)[0]
        columns.append(col)
    # <name>.md chain (BM.readBitSet :179-215)
    dbytes, pid = b"", files.get(name + ".md", INVALID_PAGE)
    while pid != INVALID_PAGE:
        base = pid * PAGE
        ln, off = struct.unpack_from(">hh", db, base + DPFIXED)
        dbytes += bytes(db[base + off:base + off + ln])
        pid = struct.unpack_from(">i", db, base + 12)[0]
    dbytes = dbytes.ljust(((len(dbytes) + 7) // 8) * 8, b"\0")
    deleted = np.frombuffer(dbytes, dtype=np.uint64).copy() if dbytes else np.zeros(0, np.uint64)
    return {"colnames": colnames, "coldescs": coldescs, "columns": columns, "deleted": deleted}


# ---- persisted bitmap indexes (reader side; the writer under test is the product's dbfile.persist_bitmap_index) --------
def read_bitmap_file(db: bytes, filename: str) -> np.ndarray:
    """bitmap/BitMapFile.java:43-60 BitMapFile(String) -> bitmap/BM.java:179-215 readBitSet: the first record of the header
    page and of every page chained behind it, concatenated, BitSet.valueOf (little-endian).  Returns uint64 words."""
    files = _file_entries(db)
    if filename not in files:
        raise KeyError(f"The file {filename} does not exist. Please provide a vaild BitMap file name.")
    data, pid = b"", files[filename]
    while pid != INVALID_PAGE:
        base = pid * PAGE
        cnt = struct.unpack_from(">h", db, base)[0]
        first = None
        for s in range(cnt):                                                  # HFPage.firstRecord: the first non-empty slot
            ln, off = struct.unpack_from(">hH", db, base + DPFIXED + 4 * s)
            if ln >= 0:
                first = (ln, off)
                break
        if first is None:
            raise ValueError(f"bitmap page {pid} of {filename} holds no record")
        data += bytes(db[base + first[1]:base + first[1] + first[0]])
        pid = struct.unpack_from(">i", db, base + 12)[0]
    data = data.rstrip(b"\0")
    data = data.ljust((len(data) + 7) // 8 * 8, b"\0")
    return np.frombuffer(data, dtype=np.uint64).copy() if data else np.zeros(0, np.uint64)


def read_bitmap_catalogue(db: bytes, name: str) -> dict:
    """columnar/Columnarfile.java:288-323: bitmapExist (the header's sixth record) and the "<col>.<value>" records behind
    it -> {"bitmapExist": [...], "values": {col: [value, ...]}} with values typed like the column."""
    files = _file_entries(db)
    recs = [r for _, _, r in _heap_records(db, files[name + ".hdr"])]
    n = struct.unpack(">i", recs[0][:4])[0]
    types = [struct.unpack_from(">i", recs[1], 4 * i)[0] for i in range(n)]
    values: dict = {c: [] for c in range(n)}
    for r in recs[6:]:
        ln = struct.unpack(">H", r[:2])[0]
        col, val = r[2:2 + ln].decode("utf-8").split(".", 1)                  # the Java splits on every dot: dotted strings break it
        values[int(col)].append(val if types[int(col)] == ATTR_STRING else int(val))
    return {"bitmapExist": list(recs[5][:n]), "values": values}


def read_bitmap_index(db: bytes, name: str, col: int) -> dict:
    """{value: BitSet words} of a persisted index: every catalogued value's file `<name>.bm.<col>.<value>`."""
    cat = read_bitmap_catalogue(db, name)
    return {v: read_bitmap_file(db, f"{name}.bm.{col}.{v}") for v in cat["values"][col]}


def space_map_pages(db: bytes) -> set:
    """Pages marked allocated in the space map (diskmgr/DB.java:739-822: bit p of the map, least significant bit first)."""
    num_pages = struct.unpack_from(">i", db, PAGE - 4)[0]
    out = set()
    for pid in range(min(num_pages, len(db) // PAGE)):
        at = PAGE * (1 + pid // (PAGE * 8)) + (pid % (PAGE * 8)) // 8
        if at < len(db) and (db[at] >> (pid % 8)) & 1:
            out.add(pid)
    return out
