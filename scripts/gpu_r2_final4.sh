#!/bin/bash
# After removing three unused helpers from the scan header: scan tests, then the traffic capture and the N=1 bench line of the
# final sources (the capture's source hash must match for bench.py to quote roofline.traffic).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_scan_gpu.py tests/test_edges_gpu.py tests/test_scan_engines_gpu.py -m gpu -x -q --timeout 600 -k "not fused and not gather" 2>&1 | tail -3
timeout -s KILL 200 python scripts/profile_scan.py 100000000 1 > gpurun_out/plain_scan.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|tile_offsets|write_kernel|write_staged|agg_finish' -o gpurun_out/scan_full_r2 python scripts/profile_scan.py 100000000 1 > gpurun_out/ncu_scan.log 2>&1
tail -1 gpurun_out/ncu_scan.log
python scripts/ncu_traffic.py gpurun_out/scan_full_r2.ncu-rep 100000000 | cut -c1-200
cp profiles/r2/traffic.json gpurun_out/traffic.json
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 200 gpurun_out/bench_n1.err; cut -c1-260 gpurun_out/bench_n1.json
