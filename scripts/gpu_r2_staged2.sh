#!/bin/bash
# write_staged_kernel: timing, then one ncu --set full capture of it at 50 % (after the plain run has exited 0).
set -x
cd /root/repo
mkdir -p gpurun_out
ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 ${SELS:-0.15,0.25,0.5,0.9} > gpurun_out/staged_eng.log 2>&1
grep -h median_ms gpurun_out/staged_eng.log | cut -c1-200
ENGINES=twopass timeout -s KILL 200 python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/plain_for_ncu.log 2>&1 && \
ENGINES=twopass timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:write_staged_kernel -s 3 -c 1 -o gpurun_out/write_staged python scripts/bench_engines.py 100000000 1 0.5 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
