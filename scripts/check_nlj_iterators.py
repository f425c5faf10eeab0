#!/usr/bin/env python
"""Runs every FILESCAN / COLUMNSCAN `nlj` command of the golden transcript through NljQuery twice -- directly on the join
kernels and through two scan iterators + iterator.ColumnarNestedLoopJoins -- and compares the printed lines."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
from mbcol.global_ import SystemDefs
from mbcol.input import NljQuery

orc.build()
names, descs, cols = orc.load_tsv(os.path.join(ROOT, "tests", "golden", "minidata.tsv"))
w = orc.DBWriter()
for cf in ("cf", "cf1", "cf2"):
    orc.write_columnar_file(w, cf, names, descs, cols)
path = os.path.join(tempfile.mkdtemp(), "db")
open(path, "wb").write(w.tobytes())
SystemDefs(path, 0, 100, None)
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "phase3_golden.json")))["entries"]
n = bad = 0
seen = set()
for e in golden:
    if e["kind"] != "nlj" or e.get("failed") or e["cmd"] in seen or "ff1." in e["cmd"] or "BTREE" in e["cmd"] or "BITMAP" in e["cmd"]:
        continue
    seen.add(e["cmd"])
    a = NljQuery().execute(e["cmd"].split()[1:], echo=False)
    try:
        b = NljQuery().execute(e["cmd"].split()[1:], echo=False, via_iterators=True)
    except Exception as ex:                                       # noqa: BLE001
        b = ["EXCEPTION " + repr(ex)]
    n += 1
    if a != b:
        bad += 1
        print("MISMATCH", e["cmd"], len(a), len(b), b[:3])
print("checked", n, "mismatches", bad)
SystemDefs.shutdown()
sys.exit(1 if bad else 0)
