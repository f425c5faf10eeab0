#!/bin/bash
# TMA-fed bitmap build for <= 32 values: bitmap tests, then the C3 builds (500 M rows) with and without it.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_bitmap_gpu.py tests/test_bitmap_persist_gpu.py -x -q --timeout 300 2>&1 | tail -4
timeout -s KILL 600 python scripts/bench_c3_c4.py --reps 5 --skip-c4 > gpurun_out/bench_c3_few.json 2> gpurun_out/bench_c3_few.err
tail -3 gpurun_out/bench_c3_few.err; cut -c1-1000 gpurun_out/bench_c3_few.json
MBC_BM_FEW_OFF=1 timeout -s KILL 600 python scripts/bench_c3_c4.py --reps 5 --skip-c4 > gpurun_out/bench_c3_fewoff.json 2> gpurun_out/bench_c3_fewoff.err
cut -c1-1000 gpurun_out/bench_c3_fewoff.json
