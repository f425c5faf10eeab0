#!/bin/bash
# Three-way write pass (group / per-tile gather / staged) and the L2 fetch granularity knob.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/split_tests.log
cat gpurun_out/split_tests.log
SELS=0.01,0.1,0.15,0.25,0.33,0.5,0.9
ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 $SELS > gpurun_out/split_default.log 2>&1
MBC_L2_FETCH=32 ENGINES=twopass,gather timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 $SELS > gpurun_out/split_l2_32.log 2>&1
MBC_L2_FETCH=128 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.01,0.1,0.5 > gpurun_out/split_l2_128.log 2>&1
grep -h median_ms gpurun_out/split_default.log gpurun_out/split_l2_32.log gpurun_out/split_l2_128.log | cut -c1-170
