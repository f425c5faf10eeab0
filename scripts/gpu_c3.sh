mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bitmap_gpu.py tests/test_mirror_gpu.py -m gpu -x -q --timeout 90 2>&1 | tail -3
for R in default 2048 4096; do
  echo "== chunk rows $R"
  if [ $R = default ]; then unset MBC_BM_CHUNK_ROWS; else export MBC_BM_CHUNK_ROWS=$R; fi
  timeout 300 python scripts/bench_c3_c4.py --skip-c4 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print({k:(round(v['ms'],2), round(v['achieved_gbs'])) for k,v in d['build'].items()}, 'scan', round(d['scan']['ms'],3))
    else: print(ln[:200])
"
done
