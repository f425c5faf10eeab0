N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?"
head -c 300 gpurun_out/bench_n$N.json; echo
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value']/1e9, 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value']/1e9)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -5
