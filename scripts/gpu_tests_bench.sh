mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python scripts/profile_scan.py 100000000 3 > gpurun_out/profile_plain.log 2>&1; tail -12 gpurun_out/profile_plain.log
