#!/bin/bash
# Programmatic dependent launch between the kernels of a scan + pipelined bench step: parity, then C2 / C5 steps with and without.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py tests/test_mirror_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/pdl_tests.log
cat gpurun_out/pdl_tests.log
for v in 1 0; do
MBC_PDL=$v timeout -s KILL 300 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/pdl${v}_c2.log 2>&1; tail -1 gpurun_out/pdl${v}_c2.log | cut -c1-420
MBC_PDL=$v timeout -s KILL 300 python bench.py --workload c5 --no-e2e --steps 20 --warmup 5 > gpurun_out/pdl${v}_c5.log 2>&1; tail -1 gpurun_out/pdl${v}_c5.log | cut -c1-420
done
MBC_PDL=1 ENGINES=twopass timeout -s KILL 400 python scripts/bench_engines.py 100000000 15 0.001,0.01,0.1,0.25,0.5,0.9 > gpurun_out/pdl_eng.log 2>&1
grep -h median_ms gpurun_out/pdl_eng.log | cut -c1-170
