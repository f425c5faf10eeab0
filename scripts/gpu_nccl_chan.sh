NG=${1:-2}
for ch in default 8 16 32; do
  if [ $ch = default ]; then unset NCCL_MIN_P2P_NCHANNELS NCCL_MAX_P2P_NCHANNELS; else export NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=$ch; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$ch', {k:d[k] for k in ('value','ms_per_step','n_gpus')})"
done
