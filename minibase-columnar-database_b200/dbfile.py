"""Host-side reader of the tiny schema part of a reference DB file image (the `.hdr` heapfile,
columnar/Columnarfile.java:257-323).  The data pages themselves are decoded on the GPU (csrc/mbc_ingest.cu)."""
from __future__ import annotations

import struct

PAGE, DPFIXED, ENTRY = 1024, 20, 56


def _file_entries(db: bytes) -> dict:
    out, pid = {}, 0
    while pid != -1:
        base = pid * PAGE
        nxt, n = struct.unpack_from(">ii", db, base)
        for i in range(n):
            o = base + 8 + i * ENTRY
            first = struct.unpack_from(">i", db, o)[0]
            if first != -1:
                ln = struct.unpack_from(">H", db, o + 4)[0]
                out[bytes(db[o + 6:o + 6 + ln]).decode("utf-8")] = first
        pid = nxt
    return out


def _records(db: bytes, first_dir: int):
    dpid = first_dir
    while dpid != -1:
        base = dpid * PAGE
        cnt = struct.unpack_from(">h", db, base)[0]
        for s in range(cnt):
            ln, off = struct.unpack_from(">hH", db, base + DPFIXED + 4 * s)
            if ln < 0:
                continue
            data_pid = struct.unpack_from(">i", db, base + off + 4)[0]
            pb = data_pid * PAGE
            for ps in range(struct.unpack_from(">h", db, pb)[0]):
                rl, ro = struct.unpack_from(">hH", db, pb + DPFIXED + 4 * ps)
                if rl >= 0:
                    yield bytes(db[pb + ro:pb + ro + rl])
        dpid = struct.unpack_from(">i", db, base + 12)[0]


def read_header(db: bytes, name: str) -> dict:
    files = _file_entries(db)
    if name + ".hdr" not in files:
        raise Exception("Columnar File does not exist.")
    recs = list(_records(db, files[name + ".hdr"]))
    n = struct.unpack(">i", recs[0][:4])[0]
    names = []
    for i in range(n):
        ln = struct.unpack_from(">H", recs[3], 17 * i)[0]
        names.append(recs[3][17 * i + 2:17 * i + 2 + ln].decode("utf-8"))
    # <name>.md : first record of every page of the chain (bitmap/BM.java:179-215)
    dbytes, pid = b"", files.get(name + ".md", -1)
    while pid != -1:
        base = pid * PAGE
        if struct.unpack_from(">h", db, base)[0] > 0:
            ln, off = struct.unpack_from(">hH", db, base + DPFIXED)
            if ln > 0:
                dbytes += bytes(db[base + off:base + off + ln])
        pid = struct.unpack_from(">i", db, base + 12)[0]
    dbytes = dbytes.rstrip(b"\0")
    dbytes = dbytes.ljust((len(dbytes) + 7) // 8 * 8, b"\0")
    return {"numColumns": n, "colnames": names, "deleted_bytes": dbytes,
            "bitmapExist": list(recs[5][:n]) if len(recs) > 5 else [0] * n,
            "bitmapValues": [r[2:2 + struct.unpack(">H", r[:2])[0]].decode("utf-8") for r in recs[6:]]}


# ---- bitmap persistence in the reference's format -----------------------------------------------
# A GPU-built bitmap index is written back into the DB file image exactly as columnar/Columnarfile.java:698-753
# (createBitMapIndex) leaves it, so that the unmodified Java `BitMapFile(String)` (bitmap/BitMapFile.java:43-60) and
# `Columnarfile(String)` (columnar/Columnarfile.java:296-323) open it:
#   * one file `<cf>.bm.<col>.<value>` per distinct value: a chain of BMIndexPages (HFPage layout, header page type BMHEAD),
#     each holding ONE record of 1000 bytes = MAX_SPACE - DPFIXED - 4 (bitmap/BM.java:64-129): BitSet.toByteArray()
#     (little-endian, trailing zero bytes dropped) cut into 1000-byte pieces, the last one zero padded (BM.java:362-374);
#   * one record `"<col>.<value>"` (Convert.setStrValue: 2-byte length + UTF) per value appended to `<cf>.hdr`
#     (Columnarfile.java:708-716 / 730-738), and bitmapExist[col] = 1 in the header's sixth record (:749-750);
#   * pages come from the space map first-fit (diskmgr/DB.java:234-327), file entries go to the first free directory slot
#     (DB.java:380-495), heapfile records to the first data page with room (heap/Heapfile.java:606-800).
INVALID_PAGE = -1
MAX_NAME = 50
BM_RECORD = PAGE - DPFIXED - 4          # 1000
NODE_BMHEAD = 13                        # btree/NodeType.java:18
DPINFO_SIZE = 8                         # heap/DataPageInfo.java:19-29: availspace:short, recct:short, pageId:int
BITS_PER_MAP_PAGE = PAGE * 8


class DBImage:
    """An editable reference DB file image (bytearray of 1024-byte pages)."""

    def __init__(self, db_bytes):
        self.b = bytearray(db_bytes)
        if len(self.b) < PAGE:
            raise ValueError("not a DB file image")
        self.num_pages = struct.unpack_from(">i", self.b, PAGE - 4)[0]          # DBFirstPage.NUM_DB_PAGE (DB.java:1000-1050)
        self.num_map_pages = (self.num_pages + BITS_PER_MAP_PAGE - 1) // BITS_PER_MAP_PAGE

    def tobytes(self) -> bytes:
        return bytes(self.b)

    def _need(self, pid: int) -> int:
        end = (pid + 1) * PAGE
        if len(self.b) < end:
            self.b.extend(bytes(end - len(self.b)))
        return pid * PAGE

    # -- diskmgr/DB.java:234-327 allocate_page (run of 1) + :739-822 set_bits ------------------------------------------
    def allocate_page(self) -> int:
        for m in range(self.num_map_pages):
            base = self._need(1 + m)
            nbits = min(BITS_PER_MAP_PAGE, self.num_pages - m * BITS_PER_MAP_PAGE)
            for byte in range((nbits + 7) // 8):
                v = self.b[base + byte]
                if v == 0xFF:
                    continue
                for bit in range(8):
                    if byte * 8 + bit < nbits and not (v >> bit) & 1:            # first fit, least significant bit first
                        self.b[base + byte] = v | (1 << bit)
                        pid = m * BITS_PER_MAP_PAGE + byte * 8 + bit
                        self._need(pid)
                        return pid
        raise Exception("No space left")                                           # OutOfSpaceException (DB.java:318)

    def page_is_allocated(self, pid: int) -> bool:
        base = (1 + pid // BITS_PER_MAP_PAGE) * PAGE
        bit = pid % BITS_PER_MAP_PAGE
        return base + bit // 8 < len(self.b) and bool((self.b[base + bit // 8] >> (bit % 8)) & 1)

    # -- diskmgr/DB.java:380-495 add_file_entry / :577-640 get_file_entry ----------------------------------------------
    def get_file_entry(self, name: str):
        return _file_entries(self.b).get(name)

    def add_file_entry(self, name: str, first_pid: int) -> None:
        if len(name) >= MAX_NAME:
            raise Exception("DB filename too long")
        if self.get_file_entry(name) is not None:
            raise Exception("DB fileentry already exists")
        raw = name.encode("utf-8")
        pid = 0
        while True:
            base = pid * PAGE
            nxt, n = struct.unpack_from(">ii", self.b, base)
            for i in range(n):
                o = base + 8 + i * ENTRY
                if struct.unpack_from(">i", self.b, o)[0] == INVALID_PAGE:
                    struct.pack_into(">iH", self.b, o, first_pid, len(raw))
                    self.b[o + 6:o + 6 + len(raw)] = raw
                    return
            if nxt == INVALID_PAGE:
                break
            pid = nxt
        new = self.allocate_page()                                                 # directory full: chain a DBDirectoryPage
        nb = self._need(new)
        self.b[nb:nb + PAGE] = bytes(PAGE)
        n = (PAGE - 16) // ENTRY                                                   # DIR_PAGE_USED_BYTES = 16
        struct.pack_into(">ii", self.b, nb, INVALID_PAGE, n)
        for i in range(n):
            struct.pack_into(">i", self.b, nb + 8 + i * ENTRY, INVALID_PAGE)
        struct.pack_into(">i", self.b, pid * PAGE, new)
        struct.pack_into(">iH", self.b, nb + 8, first_pid, len(raw))
        self.b[nb + 14:nb + 14 + len(raw)] = raw

    # -- heap/HFPage.java --------------------------------------------------------------------------------------------------
    def _hf_init(self, pid: int, page_type: int = 0) -> int:
        base = self._need(pid)
        self.b[base:base + PAGE] = bytes(PAGE)
        struct.pack_into(">hhhhiii", self.b, base, 0, PAGE, PAGE - DPFIXED, page_type, INVALID_PAGE, INVALID_PAGE, pid)
        return base

    def _hf_insert(self, pid: int, rec: bytes):
        """HFPage.insertRecord (:337-396): reuse the first empty slot, else a new one; None when the page is full."""
        base = pid * PAGE
        cnt, used, free = struct.unpack_from(">hhh", self.b, base)
        if len(rec) + 4 > free:
            return None
        slot = cnt
        for i in range(cnt):
            if struct.unpack_from(">h", self.b, base + DPFIXED + 4 * i)[0] == -1:
                slot = i
                break
        if slot == cnt:
            free -= len(rec) + 4
            cnt += 1
        else:
            free -= len(rec)
        used -= len(rec)
        struct.pack_into(">hhh", self.b, base, cnt, used, free)
        struct.pack_into(">hh", self.b, base + DPFIXED + 4 * slot, len(rec), used)
        self.b[base + used:base + used + len(rec)] = rec
        return slot

    def _hf_available(self, pid: int) -> int:
        return struct.unpack_from(">h", self.b, pid * PAGE + 4)[0] - 4            # HFPage.available_space (:621-626)

    def _hf_slots(self, pid: int):
        base = pid * PAGE
        for s in range(struct.unpack_from(">h", self.b, base)[0]):
            ln, off = struct.unpack_from(">hH", self.b, base + DPFIXED + 4 * s)
            if ln >= 0:
                yield s, ln, base + off

    # -- heap/Heapfile.java:606-800 insertRecord ---------------------------------------------------------------------------
    def heap_insert(self, first_dir: int, rec: bytes):
        dir_pid = first_dir
        while True:
            for slot, ln, at in self._hf_slots(dir_pid):                           # a data page with room, first fit
                avail, recct, data_pid = struct.unpack_from(">hhi", self.b, at)
                if len(rec) <= avail:
                    return self._heap_put(dir_pid, at, data_pid, recct, rec)
            if self._hf_available(dir_pid) >= DPINFO_SIZE:                         # a new data page, recorded on this directory page
                data_pid = self.allocate_page()
                self._hf_init(data_pid)
                info = struct.pack(">hhi", self._hf_available(data_pid), 0, data_pid)
                slot = self._hf_insert(dir_pid, info)
                if slot is None:
                    raise Exception("no space to insert rec.")
                ln, off = struct.unpack_from(">hH", self.b, dir_pid * PAGE + DPFIXED + 4 * slot)
                return self._heap_put(dir_pid, dir_pid * PAGE + off, data_pid, 0, rec)
            nxt = struct.unpack_from(">i", self.b, dir_pid * PAGE + 12)[0]
            if nxt == INVALID_PAGE:                                                # append a directory page
                nxt = self.allocate_page()
                nb = self._hf_init(nxt)
                struct.pack_into(">i", self.b, nb + 8, dir_pid)                    # prev
                struct.pack_into(">i", self.b, dir_pid * PAGE + 12, nxt)           # next
            dir_pid = nxt

    def _heap_put(self, dir_pid: int, info_at: int, data_pid: int, recct: int, rec: bytes):
        if self._hf_available(data_pid) < len(rec):
            raise Exception("no available space")                                  # SpaceNotAvailableException
        slot = self._hf_insert(data_pid, rec)
        struct.pack_into(">hhi", self.b, info_at, self._hf_available(data_pid), recct + 1, data_pid)
        return data_pid, slot

    def heap_records(self, first_dir: int):
        """(byte offset in the image, length) of every record, in Scan order."""
        dpid = first_dir
        while dpid != INVALID_PAGE:
            for _, _, at in self._hf_slots(dpid):
                data_pid = struct.unpack_from(">i", self.b, at + 4)[0]
                for _, ln, rat in self._hf_slots(data_pid):
                    yield rat, ln
            dpid = struct.unpack_from(">i", self.b, dpid * PAGE + 12)[0]

    # -- bitmap/BitMapFile.java:145-162 (BitMapFile(filename, columnfile)) + :305-310 saveToDisk -> BM.insertBitSet --------
    def write_bitset_file(self, filename: str, bitset_bytes: bytes) -> list:
        """Create the bitmap file and store the BitSet (its toByteArray() bytes); returns the page ids of the chain."""
        if self.get_file_entry(filename) is not None:
            raise Exception(f"The BitMapFile {filename} is already created.")
        head = self.allocate_page()                                                # initBitMapHeaderPage (:207-213)
        self._hf_init(head, NODE_BMHEAD)
        self.add_file_entry(filename, head)
        data = bytes(bitset_bytes).rstrip(b"\0")                                   # BitSet.toByteArray drops trailing zero bytes
        if not data:
            raise Exception("empty BitSet: BM.insertBitSet has nothing to store")  # the Java indexes bitSetBitArrays[0] (BM.java:91)
        pages, prev = [head], head
        for k in range(0, len(data), BM_RECORD):
            pid = head if k == 0 else self.allocate_page()                         # new BMIndexPage() per further piece (:102-104)
            if k:
                self._hf_init(pid)
                struct.pack_into(">i", self.b, pid * PAGE + 8, prev)               # insertRecord(rec, prevPageId, INVALID) (BMIndexPage.java:46-51)
                struct.pack_into(">i", self.b, prev * PAGE + 12, pid)              # prevBitMapPage.setNextPage (:117)
                pages.append(pid)
            if self._hf_insert(pid, data[k:k + BM_RECORD].ljust(BM_RECORD, b"\0")) is None:
                raise Exception("Insertion Record Failed")
            prev = pid
        return pages


def persist_bitmap_index(db_bytes, cf_name: str, column_no: int, values, bitsets) -> bytes:
    """columnar/Columnarfile.java:698-753 for a GPU-built index: `values[k]` (int or str) with `bitsets[k]` = the value's
    BitSet as bytes (BitSet.toByteArray() order; uint64 little-endian words serve).  Returns the new DB image."""
    img = DBImage(db_bytes)
    hdr = img.get_file_entry(cf_name + ".hdr")
    if hdr is None:
        raise Exception("Columnar File does not exist.")
    recs = list(img.heap_records(hdr))
    if len(recs) < 6:
        raise Exception(f"{cf_name}.hdr has {len(recs)} records")
    exist_at, exist_len = recs[5]                                                  # bitmapExist (Columnarfile.java:288-291)
    if column_no < 0 or column_no >= exist_len:
        raise Exception(f"Invalid column number.{column_no}")
    if img.b[exist_at + column_no] == 1:
        return img.tobytes()                                                       # `if (bitmapExist[columnNo] != 1)` (:699)
    for v, bits in zip(values, bitsets):
        name = f"{column_no}.{v}"
        img.write_bitset_file(f"{cf_name}.bm.{name}", bytes(bits))
        raw = name.encode("utf-8")
        img.heap_insert(hdr, struct.pack(">H", len(raw)) + raw)                    # Convert.setStrValue into name.length() + 2 bytes
    img.b[exist_at + column_no] = 1                                                # hdrHeapfile.updateRecord(bitmapRID, ...) (:750)
    return img.tobytes()
