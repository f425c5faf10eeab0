#!/bin/bash
set -x
mkdir -p gpurun_out
D=$PWD/minibase-columnar-database_b200/csrc/variants
MBC_LIB_PATH=$D/libmbcol_dbg.so timeout -s KILL 200 python scripts/fused_debug.py 100000000 > gpurun_out/dbg_100m.log 2>&1
tail -25 gpurun_out/dbg_100m.log | cut -c1-300
MBC_LIB_PATH=$D/libmbcol_dbg.so timeout -s KILL 200 python scripts/fused_debug.py 20000000 > gpurun_out/dbg_20m.log 2>&1
tail -8 gpurun_out/dbg_20m.log | cut -c1-300
