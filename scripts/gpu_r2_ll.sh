#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ll_c5.csv python bench.py --workload c5 --no-e2e --steps 2 --warmup 3 > gpurun_out/ll_c5.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/ll_c2.csv python bench.py --no-e2e --steps 2 --warmup 3 > gpurun_out/ll_c2.log 2>&1
tail -3 gpurun_out/ll_c2.log
