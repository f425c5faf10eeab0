#!/bin/bash
# Lean streaming write kernel: parity of the scan engines, then timing at 4/5/6 CTAs per SM, then the C5 shard size (persistent form).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -8 > gpurun_out/lean_tests.log
cat gpurun_out/lean_tests.log
V=minibase-columnar-database_b200/csrc/variants
ENGINES=twopass,stream,gather timeout -s KILL 300 python scripts/bench_engines.py 100000000 15 0.25,0.5,0.75,0.9 > gpurun_out/lean_s5.log 2>&1
MBC_LIB_PATH=$PWD/$V/libmbcol_s4.so ENGINES=stream timeout -s KILL 300 python scripts/bench_engines.py 100000000 15 0.25,0.5,0.75,0.9 > gpurun_out/lean_s4.log 2>&1
MBC_LIB_PATH=$PWD/$V/libmbcol_s6.so ENGINES=stream timeout -s KILL 300 python scripts/bench_engines.py 100000000 15 0.25,0.5,0.75,0.9 > gpurun_out/lean_s6.log 2>&1
MBC_STREAM_LEAN_OFF=1 ENGINES=stream timeout -s KILL 300 python scripts/bench_engines.py 100000000 15 0.5,0.9 > gpurun_out/lean_off.log 2>&1
grep -h median_ms gpurun_out/lean_s5.log gpurun_out/lean_s4.log gpurun_out/lean_s6.log gpurun_out/lean_off.log | cut -c1-200
timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/lean_c5.log 2>&1; tail -2 gpurun_out/lean_c5.log | cut -c1-400
