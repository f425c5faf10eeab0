mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_join_gpu.py tests/test_mirror_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -4
timeout 600 python scripts/bench_c3_c4.py --skip-c3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 exit $?"; cat gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
