#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/gpu_tests_stream.log
cat gpurun_out/gpu_tests_stream.log
ENGINES=twopass,gather timeout -s KILL 300 python scripts/bench_engines.py 100000000 10 > gpurun_out/bench_stream.log 2>&1
cat gpurun_out/bench_stream.log | cut -c1-300
