#!/usr/bin/env python
"""K1 at scale: decode a C2-shaped reference DB image (one int column of `rows` rows = rows/125 data pages of 1 KB + their
directory pages) on the GPU.  The image is built with numpy in the reference's page format (heap/HFPage.java:31-40 header,
slot directory from byte 20, records packed down from byte 1024; heap/DataPageInfo.java:19-29 directory records;
diskmgr/DB.java page 0) -- vectorised, because the oracle's row-at-a-time writer would take hours at this size; the small
files (`.hdr`, `.md`, `.dtid`) still come from the oracle's writer.  Prints one JSON line.

    python scripts/bench_ingest.py [rows] [reps]"""
import json
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol
from oracle import oracle as orc
from bench import measured_peak_gbs

PAGE, DP, PER_PAGE, PER_DIR = 1024, 20, 125, 83


def int_column_image(values: np.ndarray) -> bytes:
    n = values.size
    w = orc.DBWriter(num_pages=1 << 22)
    orc.write_columnar_file(w, "cf", ["I1"], [(1, 4)], [np.zeros(0, np.int32)])          # .hdr / empty .0 / .md / .dtid
    base = bytearray(w.tobytes())
    first_free = len(base) // PAGE
    ndata = (n + PER_PAGE - 1) // PER_PAGE
    ndir = (ndata + PER_DIR - 1) // PER_DIR
    # page ids: directory pages first, then the data pages (any order is a valid heapfile: pages are linked by id)
    dir_ids = first_free + np.arange(ndir)
    data_ids = first_free + ndir + np.arange(ndata)
    pages = np.zeros((ndir + ndata, PAGE), dtype=np.uint8)

    def be16(a):
        return np.stack([(a >> 8) & 255, a & 255], -1).astype(np.uint8)

    def be32(a):
        a = a.astype(np.int64) & 0xFFFFFFFF
        return np.stack([(a >> 24) & 255, (a >> 16) & 255, (a >> 8) & 255, a & 255], -1).astype(np.uint8)

    # ---- data pages: slotCnt, usedPtr, freeSpace, type, prev, next, cur; slots {len=4, off}; records from the end
    d = pages[ndir:]
    cnt = np.full(ndata, PER_PAGE)
    cnt[-1] = n - PER_PAGE * (ndata - 1)
    d[:, 0:2] = be16(cnt)
    d[:, 2:4] = be16(PAGE - 4 * cnt)
    d[:, 4:6] = be16(PAGE - DP - 8 * cnt)
    d[:, 8:12] = be32(np.full(ndata, -1))
    d[:, 12:16] = be32(np.full(ndata, -1))
    d[:, 16:20] = be32(data_ids)
    s = np.arange(PER_PAGE)
    d[:, DP + 4 * s] = 0
    d[:, DP + 4 * s + 1] = 4
    off = PAGE - 4 * (s + 1)
    d[:, DP + 4 * s + 2] = (off >> 8).astype(np.uint8)
    d[:, DP + 4 * s + 3] = (off & 255).astype(np.uint8)
    padded = np.zeros(ndata * PER_PAGE, dtype=np.int32)
    padded[:n] = values
    recs = be32(padded).reshape(ndata, PER_PAGE, 4)
    for k in range(4):
        d[:, PAGE - 4 * (s + 1) + k] = recs[:, :, k]
    if cnt[-1] < PER_PAGE:                                         # slots past the last record do not exist
        d[-1, DP + 4 * cnt[-1]:PAGE - 4 * cnt[-1]] = 0
    # ---- directory pages: records {availspace, recct, pageId} of 8 bytes
    r = pages[:ndir]
    dcnt = np.full(ndir, PER_DIR)
    dcnt[-1] = ndata - PER_DIR * (ndir - 1)
    r[:, 0:2] = be16(dcnt)
    r[:, 2:4] = be16(PAGE - 8 * dcnt)
    r[:, 4:6] = be16(PAGE - DP - 12 * dcnt)
    r[:, 8:12] = be32(np.concatenate([[-1], dir_ids[:-1]]))
    r[:, 12:16] = be32(np.concatenate([dir_ids[1:], [-1]]))
    r[:, 16:20] = be32(dir_ids)
    t = np.arange(PER_DIR)
    r[:, DP + 4 * t] = 0
    r[:, DP + 4 * t + 1] = 8
    doff = PAGE - 8 * (t + 1)
    r[:, DP + 4 * t + 2] = (doff >> 8).astype(np.uint8)
    r[:, DP + 4 * t + 3] = (doff & 255).astype(np.uint8)
    ids = np.full(ndir * PER_DIR, -1, dtype=np.int64)
    ids[:ndata] = data_ids
    avail = np.zeros(ndir * PER_DIR, dtype=np.int64)
    avail[:ndata] = PAGE - DP - 8 * cnt - 4
    rc = np.zeros(ndir * PER_DIR, dtype=np.int64)
    rc[:ndata] = cnt
    rec8 = np.concatenate([be16(avail), be16(rc), be32(ids)], -1).reshape(ndir, PER_DIR, 8)
    for k in range(8):
        r[:, PAGE - 8 * (t + 1) + k] = rec8[:, :, k]
    if dcnt[-1] < PER_DIR:
        r[-1, DP + 4 * dcnt[-1]:PAGE - 8 * dcnt[-1]] = 0
    # ---- point the file entry of cf.0 at the new directory chain
    o = 0
    nent = struct.unpack_from(">i", base, 4)[0]
    for i in range(nent):
        e = 8 + 56 * i
        ln = struct.unpack_from(">H", base, e + 4)[0]
        if bytes(base[e + 6:e + 6 + ln]) == b"cf.0":
            struct.pack_into(">i", base, e, int(dir_ids[0]))
    return bytes(base) + pages.tobytes()


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    vals = orc.synth_int(20260101, 0, rows, 1 << 20)
    t0 = time.perf_counter()
    img = int_column_image(vals)
    build_s = time.perf_counter() - t0
    ctx = mbcol.Context(0)
    best_wall, best_kernel = 1e9, 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        t = ctx.ingest_dbfile(img, "cf")
        wall = time.perf_counter() - t0
        best_wall, best_kernel = min(best_wall, wall), min(best_kernel, ctx.last_kernel_ms)
        got = t.read_column(0)
        assert t.nrows == rows and np.array_equal(got, vals)
        live = t.scan([], want=mbcol._native.WANT_AGG, aggs=[(0, 0)])
        assert live.count == rows                                  # no phantom / missing rows
        live.close(); t.close()
    peak, _ = measured_peak_gbs()
    pages = (rows + PER_PAGE - 1) // PER_PAGE
    page_bytes = pages * PAGE
    print(json.dumps({"config": "K1 at C2 scale: one int column", "rows": rows, "data_pages": pages, "image_bytes": len(img),
                      "decode_kernel_ms": round(best_kernel, 3), "decode_gbs_on_page_bytes": round(page_bytes / best_kernel / 1e6, 1),
                      "frac_of_measured_hbm_peak": round((page_bytes + 4 * rows) / best_kernel / 1e6 / peak, 3),
                      "ingest_wall_ms_incl_host_walk_and_upload": round(best_wall * 1e3, 1),
                      "upload_gbs_wall": round(len(img) / best_wall / 1e9, 2), "image_build_s": round(build_s, 1)}))
    ctx.close()


if __name__ == "__main__":
    main()
