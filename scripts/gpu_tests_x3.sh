mkdir -p gpurun_out
for i in 1 2 3; do timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2; done
timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
