#!/usr/bin/env python
"""ColumnarSort on the GPU (mbc_sort): the C2 table sorted by an int column, by the char(16) column and by both.
    python scripts/bench_sort.py [rows]        # default 100 000 000"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbcol
from bench import DESCS, SEED, measured_peak_gbs

N = mbcol._native
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = mbcol.Context(0)
t = ctx.create_table(DESCS, rows)
t.generate(0, 0, SEED, 1 << 20); t.generate(1, 0, SEED, 1 << 20); t.generate(2, 1, SEED); t.generate(3, 2, SEED)
peak, _ = measured_peak_gbs()
out = {"config": "sort", "rows": rows}
for name, keys, words in (("int I1", [0], 1), ("char(16) S", [3], 4), ("S then I1", [3, 0], 5)):
    ms = []
    for _ in range(3):
        r = t.sort(keys, proj=[0], want=N.WANT_POSITIONS | N.WANT_COLUMNS)
        ms.append(ctx.last_kernel_ms)
        cnt = r.count
        if _ == 2 and words == 1:                                  # sortedness of the projected key column, on the device
            import torch
            from bench import _CudaArray
            p, s = r.column_device(0)
            col = torch.as_tensor(_CudaArray(p, cnt * s), device="cuda:0").view(torch.int32)
            assert bool((col[1:] >= col[:-1]).all())
        r.close()
    m = statistics.median(ms)
    alg = rows * words * 4 * 20.0                                 # per 32-bit key word: 4 radix passes x (8 B in + 8 B out + 4 B histogram read)
    out[name] = {"ms": m, "rows_per_s": rows / (m * 1e-3), "radix_passes": 4 * words, "algorithmic_gb": alg / 1e9,
                 "achieved_gbs": alg / 1e9 / (m * 1e-3), "frac_of_measured_peak": alg / 1e9 / (m * 1e-3) / peak}
print(json.dumps(out))
t.close(); ctx.close()
