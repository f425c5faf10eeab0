#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8 > gpurun_out/gpu_tests_full.log
cat gpurun_out/gpu_tests_full.log
