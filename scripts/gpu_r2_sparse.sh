#!/bin/bash
# Sparse groups: bitmap loaded with the offsets; empty rounds of the staged kernel skipped; persistent vs one-CTA-per-tile write_kernel.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/sparse_tests.log
cat gpurun_out/sparse_tests.log
SELS=0.001,0.01,0.1,0.25,0.5,0.9
ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 $SELS > gpurun_out/sparse_eng.log 2>&1
MBC_WRITE_PERSISTENT_TILES=0 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 $SELS > gpurun_out/sparse_eng_pers.log 2>&1
grep -h median_ms gpurun_out/sparse_eng.log gpurun_out/sparse_eng_pers.log | cut -c1-170
timeout -s KILL 300 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/sparse_c2.log 2>&1; tail -1 gpurun_out/sparse_c2.log | cut -c1-600
timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/sparse_c5.log 2>&1; tail -1 gpurun_out/sparse_c5.log | cut -c1-400
MBC_WRITE_PERSISTENT_TILES=1000000 timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/sparse_c5_np.log 2>&1; tail -1 gpurun_out/sparse_c5_np.log | cut -c1-400
