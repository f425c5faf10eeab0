#!/bin/bash
# Chunked group writer + staged kernel: parity, then the crossover between the two (MBC_GROUP_MAX_PCT) and 3 vs 4 CTAs per SM.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py tests/test_edges_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/groups_tests.log
cat gpurun_out/groups_tests.log
SELS=0.001,0.01,0.1,0.15,0.25,0.33,0.5,0.9
ENGINES=twopass,gather timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 $SELS > gpurun_out/groups_default.log 2>&1
MBC_GROUP_MAX_PCT=37 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.15,0.25,0.33 > gpurun_out/groups_37.log 2>&1
MBC_GROUP_MAX_PCT=12 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.15,0.25,0.33 > gpurun_out/groups_12.log 2>&1
MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_w4.so ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.01,0.1,0.25 > gpurun_out/groups_w4.log 2>&1
grep -h median_ms gpurun_out/groups_default.log gpurun_out/groups_37.log gpurun_out/groups_12.log gpurun_out/groups_w4.log | cut -c1-170
timeout -s KILL 300 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/groups_c2.log 2>&1; tail -1 gpurun_out/groups_c2.log | cut -c1-600
timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/groups_c5.log 2>&1; tail -1 gpurun_out/groups_c5.log | cut -c1-400
