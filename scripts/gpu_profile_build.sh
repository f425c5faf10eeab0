mkdir -p gpurun_out
timeout 200 python scripts/bench_c3_c4.py --skip-c4 --rows 200000000 --reps 1 > gpurun_out/c3_small.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bitmap_build_kernel' -c 4 -f -o gpurun_out/prof_build python scripts/bench_c3_c4.py --skip-c4 --rows 200000000 --reps 1 > gpurun_out/ncu_build.log 2>&1
tail -2 gpurun_out/c3_small.json | cut -c1-600; tail -3 gpurun_out/ncu_build.log
