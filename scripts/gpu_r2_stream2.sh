#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/gpu_tests_stream.log
cat gpurun_out/gpu_tests_stream.log
ENGINES=twopass timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/plain_for_ncu.log 2>&1 && \
ENGINES=twopass timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:write_kernel -s 3 -c 1 -o gpurun_out/write_stream python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
