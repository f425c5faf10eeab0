#!/bin/bash
# TMA-staged dense write kernel: parity of every scan engine, then timing against the gather / streaming paths.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py tests/test_edges_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/staged_tests.log
cat gpurun_out/staged_tests.log
ENGINES=${ENGINES:-twopass,gather} timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.01,0.1,0.15,0.25,0.5,0.75,0.9 > gpurun_out/staged_eng.log 2>&1
grep -h median_ms gpurun_out/staged_eng.log | cut -c1-200
MBC_STAGED_STAGES=2 ENGINES=twopass timeout -s KILL 300 python scripts/bench_engines.py 100000000 15 0.25,0.5,0.9 > gpurun_out/staged_s2.log 2>&1
grep -h median_ms gpurun_out/staged_s2.log | cut -c1-200
timeout -s KILL 300 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/staged_c2.log 2>&1; tail -1 gpurun_out/staged_c2.log | cut -c1-600
