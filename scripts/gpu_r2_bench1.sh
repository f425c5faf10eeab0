#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/engines_test4.log
cat gpurun_out/engines_test4.log
ENGINES=twopass,stream timeout -s KILL 300 python scripts/bench_engines.py 100000000 10 > gpurun_out/bench_stream2.log 2>&1
tail -1 gpurun_out/bench_stream2.log | cut -c1-500
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | cut -c1-3000
