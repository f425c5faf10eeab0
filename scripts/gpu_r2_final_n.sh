#!/bin/bash
# usage: gpu_r2_final_n.sh N  -- the driver's command line for N GPUs, NCCL_DEBUG=INFO like a rank check would set it
N=$1
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_n$N.json) nranks_lines=$(grep -c 'nranks' gpurun_out/bench_n$N.err)"
grep -m2 "nranks" gpurun_out/bench_n$N.err | cut -c1-200
grep -E "Traceback|Error" -A4 gpurun_out/bench_n$N.err | head -20
cut -c1-600 gpurun_out/bench_n$N.json
