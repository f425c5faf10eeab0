mkdir -p gpurun_out
timeout 300 python scripts/bench_c3_c4.py --skip-c3 --reps 1 > gpurun_out/c4_plain.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'join_unique_pass_kernel|join_agg_pass_kernel' -c 2 -f -o gpurun_out/prof_join python scripts/bench_c3_c4.py --skip-c3 --reps 1 > gpurun_out/ncu_join.log 2>&1
tail -1 gpurun_out/c4_plain.json | cut -c1-200; tail -2 gpurun_out/ncu_join.log | cut -c1-200
