#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_bitmap_gpu.py tests/test_join_gpu.py tests/test_bitmap_persist_gpu.py -x -q --timeout 300 2>&1 | tail -4
timeout -s KILL 900 python scripts/bench_c3_c4.py --reps 5 > gpurun_out/bench_c3_c4_r2.json 2> gpurun_out/bench_c3_c4_r2.err
tail -5 gpurun_out/bench_c3_c4_r2.err; cut -c1-900 gpurun_out/bench_c3_c4_r2.json
