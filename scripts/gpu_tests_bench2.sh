mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value', d['value']/1e9, 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value']/1e9, 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
PY
tail -3 gpurun_out/bench_n1.err
