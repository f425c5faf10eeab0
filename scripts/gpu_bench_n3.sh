#!/bin/bash
# bench.py + config C5 on the N GPUs of this box (N = $1)
NG=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_n$NG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 scripts/bench_c5.py > gpurun_out/c5_n$NG.json 2> gpurun_out/c5_n$NG.err; echo "c5 rc=$?"
cat gpurun_out/c5_n$NG.json
