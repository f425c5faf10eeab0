#!/bin/bash
# Round-end evidence on ONE GPU: tests, the default bench line, the ncu launch list of the same command, one full ncu
# capture of the scan kernels and of the bitmap build / join kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_short.json 2>/dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "launchlist rc=$?"
timeout 120 python scripts/profile_scan.py 100000000 2 > gpurun_out/profile_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|write_kernel|tile_offsets_kernel|agg_finish_kernel' -s 12 -c 12 -f -o gpurun_out/prof_scan_final python scripts/profile_scan.py 100000000 2 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/profile_plain.log
