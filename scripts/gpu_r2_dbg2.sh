#!/bin/bash
mkdir -p gpurun_out
D=$PWD/minibase-columnar-database_b200/csrc/variants
for P in 2 3 4 5 6; do
  echo "== P=$P"
  MBC_FUSED_PRED_STAGES=$P MBC_LIB_PATH=$D/libmbcol_dbg.so timeout -s KILL 100 python scripts/fused_debug.py 20000000 2>&1 | grep -E "^case|count|fused check|all cases" | head -20
done
echo "== P=5 S=2 g1"
MBC_FUSED_PRED_STAGES=5 MBC_FUSED_PAY_STAGES=2 MBC_LIB_PATH=$D/libmbcol_g1.so timeout -s KILL 100 python scripts/fused_debug.py 20000000 2>&1 | grep -E "^case|count|fused check|all cases|Error" | head -20
