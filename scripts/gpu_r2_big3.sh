#!/bin/bash
# Why is the persistent write_kernel slow on mid-density tiles?  ncu --set full of it at 250 M rows, 25 %.
set -x
cd /root/repo
mkdir -p gpurun_out
ENGINES=twopass timeout -s KILL 300 python scripts/bench_engines.py 250000000 1 0.25 > gpurun_out/big3_plain.log 2>&1 && \
ENGINES=twopass timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:write_kernel -s 3 -c 1 -o gpurun_out/write_pers python scripts/bench_engines.py 250000000 1 0.25 > gpurun_out/big3_ncu.log 2>&1
tail -2 gpurun_out/big3_ncu.log
