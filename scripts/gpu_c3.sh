mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bitmap_gpu.py tests/test_mirror_gpu.py -m gpu -x -q --timeout 90 2>&1 | tail -4
timeout 600 python scripts/bench_c3_c4.py --skip-c4 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 exit $?"; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
