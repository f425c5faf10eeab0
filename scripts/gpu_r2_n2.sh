#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout -s KILL 600 python -m pytest tests/test_shard_gpu.py -x -q --timeout 300 2>&1 | tail -5 > gpurun_out/shard_test_n2.log
cat gpurun_out/shard_test_n2.log
NCCL_DEBUG=INFO timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"
wc -l gpurun_out/bench_n2.json
grep -c "NCCL INFO" gpurun_out/bench_n2.err
grep -E "nranks|Error|error|Traceback" gpurun_out/bench_n2.err | head -10
tail -c 1500 gpurun_out/bench_n2.err
