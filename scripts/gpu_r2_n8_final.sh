#!/bin/bash
# N=8 on the final tree: the driver's command line for the C5 bench (full line, parity checked on rank 0).
set -x
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_c5.json 2> gpurun_out/bench_n8.err
tail -c 300 gpurun_out/bench_n8.err; cut -c1-300 gpurun_out/bench_n8_c5.json
