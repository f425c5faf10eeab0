#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_shard_gpu.py -x -q --timeout 200 2>&1 | tail -3
for v in "MBC_BENCH_GATHER=stores" "MBC_BENCH_GATHER=dma"; do
  echo "== $v"
  env $v timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e 2>gpurun_out/n2.err | cut -c1-500
done
tail -c 400 gpurun_out/n2.err
