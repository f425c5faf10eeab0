mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 120 python scripts/profile_scan.py 100000000 3 > gpurun_out/profile_plain.log 2>&1; tail -3 gpurun_out/profile_plain.log
for v in "$@"; do
  echo "== variant $v"
  MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_$v.so timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
done
