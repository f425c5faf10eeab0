#!/usr/bin/env python
"""Per-selectivity wall time of mbc_scan_host over pinned C2 columns (tuning probe for the late-materialisation threshold)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mbcol
from bench import AGGS, DESCS, SEED, c2_terms
N = mbcol._native
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
sels = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.01, 0.03, 0.1, 0.5]
ctx = mbcol.Context(0)
t = ctx.create_table(DESCS, rows)
t.generate(0, 0, SEED, 1 << 20); t.generate(1, 0, SEED, 1 << 20); t.generate(2, 1, SEED); t.generate(3, 2, SEED)
host = []
for c, (ty, w) in enumerate(DESCS):
    a = t.read_column(c)
    buf = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(a.dtype).reshape(a.shape)
    buf[...] = a
    host.append(buf)
t.close()
want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_HOST
for s in sels:
    for rep in range(3):
        b0 = ctx.h2d_bytes
        t0 = time.perf_counter()
        r = ctx.scan_host(DESCS, host, c2_terms(mbcol.Term, s), proj=[0, 1, 2, 3], want=want, aggs=AGGS)
        dt = time.perf_counter() - t0
        cnt = r.count
        r.close()
    print(f"sel {s}: count {cnt} wall_ms {1e3 * dt:.2f} h2d_MB {(ctx.h2d_bytes - b0) / 1e6:.0f} div {os.environ.get('MBC_LATE_DIV', '32')}")
ctx.close()
