mkdir -p gpurun_out
timeout 120 python scripts/profile_scan.py 100000000 3 > gpurun_out/profile_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'write_kernel' -s 3 -c 3 -f -o gpurun_out/prof_write python scripts/profile_scan.py 100000000 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/profile_plain2.log; tail -3 gpurun_out/ncu_full.log
