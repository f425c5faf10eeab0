#!/usr/bin/env python
"""Bisect a failing fused scan: runs cases one by one on a C2 table of `rows` rows, printing before each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol
from bench import AGGS, DESCS, SEED, c2_terms
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
N = mbcol._native
ctx = mbcol.Context(0)
t = ctx.create_table(DESCS, rows)
for c, (k, d) in enumerate([(0, 1 << 20), (0, 1 << 20), (1, 0), (2, 0)]):
    t.generate(c, k, SEED, d)
os.environ["MBC_SCAN_PATH"] = "fused"
cases = [(0.001, [0, 1, 2, 3], AGGS, 0), (0.001, [0, 3], [], N.WANT_HOST), (0.001, [0, 3], AGGS, 0), (0.001, [0, 1, 2, 3], [], 0),
         (0.5, [0, 3], [], N.WANT_HOST), (0.5, [0, 1, 2, 3], AGGS, 0), (0.1, [0, 3], [], 0)]
for sel, proj, aggs, extra in cases:
    print("case", sel, proj, len(aggs), extra, flush=True)
    r = t.scan(c2_terms(mbcol.Term, sel), proj=proj, want=N.WANT_POSITIONS | N.WANT_COLUMNS | (N.WANT_AGG if aggs else 0) | extra, aggs=aggs)
    print("  count", r.count, "ms", round(r.kernel_ms, 3), flush=True)
    if extra:
        pos = r.positions()
        print("  sorted", bool(np.all(np.diff(pos) > 0)), flush=True)
    r.close()
print("all cases ran", flush=True)
