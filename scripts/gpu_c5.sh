#!/bin/bash
# edge tests + config C5 on the GPUs of this box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_edges_gpu.py -x -q -m gpu --timeout 90 > gpurun_out/edges.log 2>&1; echo "edges rc=$?"
tail -15 gpurun_out/edges.log
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 scripts/bench_c5.py > gpurun_out/c5_n$NG.json 2> gpurun_out/c5_n$NG.err; echo "c5 rc=$?"
cat gpurun_out/c5_n$NG.json; tail -5 gpurun_out/c5_n$NG.err
