#!/usr/bin/env python
"""Both scan engines on the C2 table (device-resident), per selectivity: median device time of `reps` scans, §8d roofline
fraction, and agreement of counts / aggregates / a checksum of positions between the engines.

    python scripts/bench_engines.py [rows] [reps] [sel,sel,...]
Environment of the library applies (MBC_FUSED_DENSE_MIN, MBC_FUSED_PRED_STAGES, MBC_FUSED_PAY_STAGES).
Prints one JSON line per (engine, selectivity) and a final summary line."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol
from bench import AGGS, DESCS, SEED, algorithmic_bytes, c2_terms, measured_peak_gbs

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sels = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.001, 0.01, 0.03, 0.1, 0.25, 0.5, 0.9]
engines = os.environ.get("ENGINES", "twopass,gather,fused").split(",")
N = mbcol._native
ctx = mbcol.Context(0)
t = ctx.create_table(DESCS, rows)
t.generate(0, 0, SEED, 1 << 20)
t.generate(1, 0, SEED, 1 << 20)
t.generate(2, 1, SEED)
t.generate(3, 2, SEED)
peak, _ = measured_peak_gbs()
want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG
summary = {}
sigs = {}
for eng in engines:
    os.environ["MBC_SCAN_PATH"] = "fused" if eng == "fused" else "twopass"
    os.environ["MBC_WRITE_STAGED"] = "0" if eng == "gather" else "1"   # gather: the round-1 write pass (no write_staged_kernel)
    for s in sels:
        terms = c2_terms(mbcol.Term, s)
        ms = []
        for rep in range(reps + 3):
            r = t.scan(terms, proj=[0, 1, 2, 3], want=want, aggs=AGGS)
            k = r.kernel_ms
            if rep >= 3:
                ms.append(k)
            sig = (r.count, tuple((lambda g: g[0] if float(g[0]) == g[1] else float("%.9g" % g[1]))(r.agg(a)) for a in range(len(AGGS))))
            r.close()
        # one checked run: positions + a column brought to the host
        r = t.scan(terms, proj=[0, 3], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_HOST)
        pos = r.positions()
        chk = (int(pos.sum()), int(r.column(0).astype(np.int64).sum()), int(r.column(1).astype(np.uint64).sum()),
               bool(np.all(np.diff(pos) > 0)))
        r.close()
        sigs.setdefault(s, {})[eng] = (sig, chk)
        m = statistics.median(ms)
        b = algorithmic_bytes(rows, s)
        line = {"engine": eng, "sel": s, "count": sig[0], "median_ms": round(m, 4), "min_ms": round(min(ms), 4),
                "achieved_gbs": round(b / m / 1e6, 1), "frac_measured": round(b / m / 1e6 / peak, 3), "frac_8000": round(b / m / 1e6 / 8000, 3)}
        print(json.dumps(line), flush=True)
        summary.setdefault(eng, {})[s] = m
agree = all(len(set(map(repr, v.values()))) == 1 for v in sigs.values())
print(json.dumps({"rows": rows, "reps": reps, "engines_agree": agree,
                  "ms": {e: {str(s): round(v, 4) for s, v in d.items()} for e, d in summary.items()},
                  "env": {k: v for k, v in os.environ.items() if k.startswith("MBC_FUSED")}}), flush=True)
if not agree:
    for s, v in sigs.items():
        if len(set(map(repr, v.values()))) != 1:
            print("DISAGREE", s, v, flush=True)
    sys.exit(1)
t.close()
ctx.close()
