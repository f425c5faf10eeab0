mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_bitmap_gpu.py tests/test_join_gpu.py tests/test_mirror_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 600 python scripts/bench_c3_c4.py > gpurun_out/bench_c3_c4.json 2> gpurun_out/bench_c3_c4.err; echo "rc=$?"; cat gpurun_out/bench_c3_c4.json | cut -c1-3000; tail -3 gpurun_out/bench_c3_c4.err
