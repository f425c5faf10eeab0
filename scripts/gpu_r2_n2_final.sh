#!/bin/bash
# N=2 on the final tree: the shard tests (2 GPUs), then the driver's command line for the C5 bench (full line, parity checked on rank 0).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_shard_gpu.py -x -q --timeout 200 2>&1 | tail -3
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_c5.json 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2_c5.json
