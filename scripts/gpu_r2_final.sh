#!/bin/bash
# Round-2 evidence on the final tree: every GPU test, smoke, the N=1 bench line, ncu --set full of one pass of the three C2 scans
# (traffic.json), the launch list of the bench command.  Each ncu run follows its own plain run.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8 > gpurun_out/gpu_tests_full.log
cat gpurun_out/gpu_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
ENGINES=twopass,gather,fused timeout -s KILL 400 python scripts/bench_engines.py 100000000 10 > gpurun_out/bench_engines_final.log 2>&1
tail -1 gpurun_out/bench_engines_final.log | cut -c1-700
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout -s KILL 200 python scripts/profile_scan.py 100000000 1 > gpurun_out/plain_scan.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|tile_offsets|write_kernel|write_staged|agg_finish' -o gpurun_out/scan_full_r2 python scripts/profile_scan.py 100000000 1 > gpurun_out/ncu_scan.log 2>&1
tail -2 gpurun_out/ncu_scan.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pageable > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
