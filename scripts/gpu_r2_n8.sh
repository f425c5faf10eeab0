#!/bin/bash
mkdir -p gpurun_out
for v in "MBC_BENCH_GATHER=dma" "MBC_BENCH_GATHER=stores"; do
  echo "== $v"
  env $v timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e 2>gpurun_out/n8.err | cut -c1-700
done
tail -c 200 gpurun_out/n8.err
