#!/bin/bash
# Round-2 evidence, final tree (after the persistent-form routing and the 1-group default of the fused engine): every GPU test,
# smoke, engines at 100 M and 250 M rows, ncu --set full of one pass of the three C2 scans (traffic.json), the launch list, and the
# N=1 bench line LAST (so that it finds the traffic capture of its own sources only if this script's post-processing ran before).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8 > gpurun_out/gpu_tests_full.log
cat gpurun_out/gpu_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
ENGINES=twopass,gather,fused timeout -s KILL 400 python scripts/bench_engines.py 100000000 10 > gpurun_out/bench_engines_final.log 2>&1
tail -1 gpurun_out/bench_engines_final.log | cut -c1-700
ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 250000000 8 0.01,0.1,0.15,0.25,0.5 > gpurun_out/bench_engines_250m.log 2>&1
tail -1 gpurun_out/bench_engines_250m.log | cut -c1-300
timeout -s KILL 200 python scripts/profile_scan.py 100000000 1 > gpurun_out/plain_scan.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|tile_offsets|write_kernel|write_staged|agg_finish' -o gpurun_out/scan_full_r2 python scripts/profile_scan.py 100000000 1 > gpurun_out/ncu_scan.log 2>&1
tail -2 gpurun_out/ncu_scan.log
python scripts/ncu_traffic.py gpurun_out/scan_full_r2.ncu-rep 100000000 | cut -c1-300
cp profiles/r2/traffic.json gpurun_out/traffic.json
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 300 gpurun_out/bench_n1.err; cut -c1-260 gpurun_out/bench_n1.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pageable > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
