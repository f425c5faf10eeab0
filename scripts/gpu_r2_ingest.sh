#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ingest_gpu.py tests/test_mirror_gpu.py tests/test_bitmap_persist_gpu.py tests/test_edges_gpu.py -x -q --timeout 300 2>&1 | tail -6
