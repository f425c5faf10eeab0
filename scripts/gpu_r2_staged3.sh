#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_engines_gpu.py tests/test_scan_gpu.py -m gpu -x -q --timeout 600 -k "twopass or staged2 or not engine" 2>&1 | tail -12 > gpurun_out/staged_tests.log
cat gpurun_out/staged_tests.log
bash scripts/gpu_r2_staged2.sh
