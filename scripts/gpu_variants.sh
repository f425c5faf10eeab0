mkdir -p gpurun_out
for v in "$@"; do
  echo "== variant $v"
  MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_$v.so python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
done
