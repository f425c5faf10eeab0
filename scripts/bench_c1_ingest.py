#!/usr/bin/env python
"""BASELINE config C1 shape at a size worth timing: a reference-format DB file image (1 KB heap pages: page directory,
slot directories, big-endian fields) of a 4-column table is decoded into resident columns by K1 (decode_pages_kernel),
then one single-column predicate scan runs over it.  The image is built on the host by the oracle's page writer
(test infrastructure; row-at-a-time like the Java, so the row count is kept modest).

    python scripts/bench_c1_ingest.py [rows]        # default 1 000 000 rows
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import mbcol
from oracle import oracle as orc
from util import C2_DESCS, C2_NAMES, c2_columns

N = mbcol._native
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
orc.build()
cols = c2_columns(orc, rows)
t0 = time.perf_counter()
db = orc.DBWriter(num_pages=1 << 20)
orc.write_columnar_file(db, "cf", C2_NAMES, C2_DESCS, cols)
img = db.tobytes() if hasattr(db, "tobytes") else db.image()
write_s = time.perf_counter() - t0
ctx = mbcol.Context(0)
times = []
for _ in range(5):
    t1 = time.perf_counter()
    tab = ctx.ingest_dbfile(img, "cf")
    wall = time.perf_counter() - t1
    times.append((ctx.last_kernel_ms, wall * 1e3))
    if _ < 4:
        tab.close()
for c in range(4):
    assert np.array_equal(np.asarray(tab.read_column(c)).view(np.uint8), np.asarray(cols[c]).view(np.uint8)), c
terms = [mbcol.Term(N.OP_LT, ("col", 0), ("int", 1 << 19), 0)]
exp = orc.scan(C2_DESCS, cols, terms, proj=[0, 3], aggs=[(0, 0)])
res = tab.scan(terms, proj=[0, 3], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_HOST, aggs=[(0, 0)])
assert res.count == exp["count"] and np.array_equal(res.positions(), exp["positions"])
kms = sorted(t[0] for t in times)[len(times) // 2]
wms = sorted(t[1] for t in times)[len(times) // 2]
data_pages = sum((rows * w + (1004 // (w + 4)) * w - 1) // ((1004 // (w + 4)) * w) for _, w in C2_DESCS)
print(json.dumps({"config": "C1-shape ingest", "rows": rows, "image_bytes": len(img), "data_pages": int(data_pages),
                  "host_page_writer_s": write_s, "decode_kernel_ms": kms, "ingest_call_wall_ms": wms,
                  "decoded_rows_per_s": rows / (kms * 1e-3), "page_bytes_per_s_gb": data_pages * 1024 / (kms * 1e-3) / 1e9,
                  "scan_after_ingest_count": res.count}))
