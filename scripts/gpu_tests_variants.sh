mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "== default"; timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
timeout 300 python scripts/bench_c3_c4.py --skip-c4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('c3 scan ms', d['scan']['ms'])"
for v in "$@"; do
  echo "== variant $v"
  export MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_$v.so
  timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
  timeout 300 python scripts/bench_c3_c4.py --skip-c4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('c3 scan ms', d['scan']['ms'])"
done
