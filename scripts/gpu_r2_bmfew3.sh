#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_bitmap_gpu.py tests/test_bitmap_persist_gpu.py -x -q --timeout 200 2>&1 | tail -2
timeout -s KILL 300 python scripts/bench_c3_c4.py --reps 5 --skip-c4 > gpurun_out/bench_c3_minmax32.json 2> gpurun_out/bench_c3_minmax32.err
cut -c1-640 gpurun_out/bench_c3_minmax32.json
