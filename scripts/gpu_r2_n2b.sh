#!/bin/bash
mkdir -p gpurun_out
for v in "MBC_X=1" "MBC_SHARD_PUSH_CTAS=16" "MBC_SHARD_PUSH_CTAS=32" "MBC_SHARD_PUSH_CTAS=148"; do
  echo "== $v"
  env $v timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e 2>/dev/null | cut -c1-400
done
