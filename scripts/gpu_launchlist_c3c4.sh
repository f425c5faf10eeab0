mkdir -p gpurun_out
timeout 300 python scripts/bench_c3_c4.py --reps 1 > gpurun_out/c3c4_plain.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_c3c4.csv python scripts/bench_c3_c4.py --reps 1 > gpurun_out/ncu_c3c4.log 2>&1
echo rc=$?; tail -2 gpurun_out/c3c4_plain.json | cut -c1-200
