#!/bin/bash
# round 2: where does the fused kernel wait?  per-role cycle counters, ring-depth variants, one ncu --set full capture
set -x
mkdir -p gpurun_out
export ENGINES=fused
MBC_FUSED_PROF=1 timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.001,0.01,0.1,0.5 > gpurun_out/prof_default.log 2>&1
grep -E "fused prof|engine" gpurun_out/prof_default.log | tail -24
for v in "MBC_FUSED_PAY_STAGES=3 MBC_FUSED_PRED_STAGES=2" "MBC_FUSED_PRED_STAGES=2" "MBC_FUSED_DENSE_MIN=1"; do
  env $v timeout -s KILL 200 python scripts/bench_engines.py 100000000 5 0.001,0.5 > "gpurun_out/variant_$(echo $v | tr ' =' '__').log" 2>&1
  tail -1 "gpurun_out/variant_$(echo $v | tr ' =' '__').log"
done
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fused_scan -s 3 -c 1 -o gpurun_out/fused_r2a python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
