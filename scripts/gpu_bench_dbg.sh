NG=${1:-8}
mkdir -p gpurun_out
MBC_BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_dbg_n$NG.json 2> gpurun_out/bench_dbg_n$NG.err; echo "rc=$?"
grep "bench debug" gpurun_out/bench_dbg_n$NG.err; tail -1 gpurun_out/bench_dbg_n$NG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])"; nproc
