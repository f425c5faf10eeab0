#!/bin/bash
# round 2: unified fused kernel (v2): parity, per-stage cycle counters, ring-depth / look-ahead variants
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py -x -q --timeout 300 2>&1 | tail -15 > gpurun_out/engines_test_v2.log
cat gpurun_out/engines_test_v2.log
export ENGINES=fused
MBC_FUSED_PROF=1 timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.001,0.01,0.1,0.5 > gpurun_out/prof_v2.log 2>&1
grep -E "fused prof|engine" gpurun_out/prof_v2.log | tail -30
i=0
for v in "MBC_X=0" "MBC_FUSED_PAY_STAGES=3 MBC_FUSED_PRED_STAGES=2" "MBC_FUSED_PAY_STAGES=3 MBC_FUSED_PRED_STAGES=2 MBC_FUSED_AHEAD=1" "MBC_FUSED_AHEAD=1" "MBC_FUSED_AHEAD=3" "MBC_FUSED_PAY_STAGES=1" "MBC_FUSED_DENSE_MIN=16" "MBC_FUSED_DENSE_MIN=256"; do
  i=$((i+1))
  env $v timeout -s KILL 200 python scripts/bench_engines.py 100000000 7 0.001,0.01,0.03,0.1,0.25,0.5,0.9 > "gpurun_out/v2_variant_$i.log" 2>&1
  echo "$v"; tail -1 "gpurun_out/v2_variant_$i.log"
done
