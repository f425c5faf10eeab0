#!/bin/bash
# Persistent vs one-CTA-per-tile write_kernel on the mostly-empty big tables (C5 shard: 500 M rows at 1 %; 500 M at 0.01 % and 0.1 %).
set -x
cd /root/repo
mkdir -p gpurun_out
for v in 49152 100000000; do
MBC_WRITE_PERSISTENT_TILES=$v timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/big2_c5_$v.log 2>&1; tail -1 gpurun_out/big2_c5_$v.log | cut -c1-330
MBC_WRITE_PERSISTENT_TILES=$v ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 500000000 8 0.0001,0.001,0.01 > gpurun_out/big2_eng_$v.log 2>&1
grep -h median_ms gpurun_out/big2_eng_$v.log | cut -c1-120
done
