#!/bin/bash
# round 2, first GPU call: parity of the fused engine, then the engine comparison on the full C2 table
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py -x -q --timeout 300 2>&1 | tail -30 > gpurun_out/engines_test.log
cat gpurun_out/engines_test.log
timeout -s KILL 300 python scripts/bench_engines.py 100000000 10 > gpurun_out/bench_engines.log 2>&1
cat gpurun_out/bench_engines.log
