for v in default "$@"; do
  echo "== $v"
  if [ $v = default ]; then unset MBC_LIB_PATH; else export MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_$v.so; fi
  timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
  timeout 120 python scripts/profile_scan.py 100000000 3 single 2>&1 | tail -3
done
