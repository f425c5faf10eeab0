#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py -x -q --timeout 300 2>&1 | tail -15 > gpurun_out/engines_test_v3.log
cat gpurun_out/engines_test_v3.log
export ENGINES=fused
MBC_FUSED_PROF=1 timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.001,0.01,0.1,0.5 > gpurun_out/prof_v3.log 2>&1
grep -E "fused prof" gpurun_out/prof_v3.log | awk 'NR%5==1' | cut -c1-700
i=0
for v in "MBC_X=0" "MBC_FUSED_PAY_STAGES=2" "MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_g1.so" "MBC_FUSED_DENSE_MIN=16"; do
  i=$((i+1))
  env $v timeout -s KILL 200 python scripts/bench_engines.py 100000000 7 0.001,0.01,0.03,0.1,0.25,0.5,0.9 > "gpurun_out/v3_variant_$i.log" 2>&1
  echo "$v"; tail -1 "gpurun_out/v3_variant_$i.log"
done
MBC_FUSED_PROF=1 MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_g1.so timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.01,0.5 > gpurun_out/prof_v3_g1.log 2>&1
grep -E "fused prof" gpurun_out/prof_v3_g1.log | awk 'NR%5==1' | cut -c1-700
timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fused_scan -s 3 -c 1 -o gpurun_out/fused_v3 python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
