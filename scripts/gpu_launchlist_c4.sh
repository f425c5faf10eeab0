mkdir -p gpurun_out
timeout 300 python scripts/bench_c3_c4.py --skip-c3 --reps 1 > gpurun_out/c4_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python scripts/bench_c3_c4.py --skip-c3 --reps 1 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/c4_plain.json | cut -c1-300; tail -2 gpurun_out/ncu_c4.log | cut -c1-200
