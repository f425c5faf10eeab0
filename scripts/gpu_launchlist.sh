mkdir -p gpurun_out
timeout 120 python scripts/profile_scan.py 100000000 2 > gpurun_out/profile_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_scan.csv python scripts/profile_scan.py 100000000 2 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
