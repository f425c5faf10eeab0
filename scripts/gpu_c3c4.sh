mkdir -p gpurun_out
timeout 600 python scripts/bench_c3_c4.py > gpurun_out/bench_c3_c4.json 2> gpurun_out/bench_c3_c4.err; echo "rc=$?"; cat gpurun_out/bench_c3_c4.json | cut -c1-3000; tail -3 gpurun_out/bench_c3_c4.err
