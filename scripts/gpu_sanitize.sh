mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_scan_gpu.py tests/test_bitmap_gpu.py tests/test_join_gpu.py -m gpu -x -q --timeout 600 -k "not full_size and not streaming and not c2_shape" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
tail -25 gpurun_out/sanitize_memcheck.log | cut -c1-300
