#!/bin/bash
# Last check of the round: every GPU test and smoke on the final tree, then the C3 / C4 numbers (bitmap few-values kernel changed).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6 > gpurun_out/gpu_tests_full.log
cat gpurun_out/gpu_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -s KILL 900 python scripts/bench_c3_c4.py --reps 5 > gpurun_out/bench_c3_c4_r2.json 2> gpurun_out/bench_c3_c4_r2.err
tail -2 gpurun_out/bench_c3_c4_r2.err; cut -c1-400 gpurun_out/bench_c3_c4_r2.json
