#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:'bitmap_build_few|column_range' -c 4 -o gpurun_out/bmfew python scripts/bench_c3_c4.py --reps 1 --skip-c4 > gpurun_out/bmfew_ncu.log 2>&1
tail -2 gpurun_out/bmfew_ncu.log | cut -c1-200
