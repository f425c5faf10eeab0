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
