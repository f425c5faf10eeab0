mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_scan_gpu.py -x -q -m gpu --timeout 120 -k "scan_host" > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_host.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_late.json 2> gpurun_out/bench_late.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_late.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")}); print(d["roofline"]["frac"], {k:v["kernel_ms"] for k,v in d["roofline"]["per_selectivity"].items()})
P
MBC_NO_LATE=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-late e2e', d['e2e'])"
