#!/bin/bash
# Tables above 49 152 tiles use the persistent write_kernel: is its per-tile gather of mid-density groups worth keeping, or should
# every non-sparse group go to write_staged_kernel there?  (250 M C2 rows, 15 / 25 %)
set -x
cd /root/repo
mkdir -p gpurun_out
ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 250000000 8 0.01,0.1,0.15,0.25,0.5 > gpurun_out/big_default.log 2>&1
MBC_STAGED_MIN_PCT=0 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 250000000 8 0.15,0.25 > gpurun_out/big_staged.log 2>&1
MBC_WRITE_PERSISTENT_TILES=100000000 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 250000000 8 0.01,0.1,0.15,0.25,0.5 > gpurun_out/big_nonpers.log 2>&1
grep -h median_ms gpurun_out/big_default.log gpurun_out/big_staged.log gpurun_out/big_nonpers.log | cut -c1-150
