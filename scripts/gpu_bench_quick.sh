mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_edges_gpu.py tests/test_scan_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"])
for k,v in d["roofline"]["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
P
