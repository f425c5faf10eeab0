#!/usr/bin/env python
"""Small driver for ncu: builds the C2 table on cuda:0 and runs the three BASELINE scans `reps` times.
Usage: python scripts/profile_scan.py [rows] [reps] [workload]   workload in {scan, bitmap, join}"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol
from bench import AGGS, DESCS, SEED, SELECTIVITIES, c2_terms

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = mbcol._native
ctx = mbcol.Context(0)
t = ctx.create_table(DESCS, rows)
t.generate(0, 0, SEED, 1 << 20)
t.generate(1, 0, SEED, 1 << 20)
t.generate(2, 1, SEED)
t.generate(3, 2, SEED)
want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG
if len(sys.argv) > 3 and sys.argv[3] == "single":          # one predicate column, positions only (the reference's C1 query shape)
    for rep in range(reps):
        for s in SELECTIVITIES:
            r = t.scan([mbcol.Term(N.OP_LT, ("col", 0), ("int", int(s * (1 << 20))), 0)], proj=[], want=N.WANT_POSITIONS | N.WANT_AGG, aggs=[(0, 0)])
            ph = r.phase_ms
            print(f"single rep {rep} sel {s}: count {r.count} kernel_ms {r.kernel_ms:.3f} filter {ph[0]:.3f} write {ph[2]:.3f}")
            r.close()
    t.close(); ctx.close(); sys.exit(0)
for rep in range(reps):
    for s in SELECTIVITIES:
        r = t.scan(c2_terms(mbcol.Term, s), proj=[0, 1, 2, 3], want=want, aggs=AGGS)
        print(f"rep {rep} sel {s}: count {r.count} kernel_ms {ctx.last_kernel_ms:.3f}")
        r.close()
t.close()
ctx.close()
