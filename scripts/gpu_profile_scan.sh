mkdir -p gpurun_out
timeout 120 python scripts/profile_scan.py 100000000 3 > gpurun_out/profile_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_scan.csv python scripts/profile_scan.py 100000000 3 > gpurun_out/ncu_launch.log 2>&1
timeout 120 python scripts/profile_scan.py 100000000 3 > gpurun_out/profile_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|write_kernel' -s 4 -c 6 -o gpurun_out/prof_scan python scripts/profile_scan.py 100000000 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/profile_plain.log | tail -3; tail -3 gpurun_out/ncu_full.log
