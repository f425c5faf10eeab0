#!/usr/bin/env python
"""BASELINE config C5: 8-column synthetic table (I1..I5 int, R1,R2 real, S char(16) = 44 B/row) sharded by TID range,
`{(I1,<,t1)}^{(R1,<,t2)}` at 1 % selectivity, project [I1,I2,R1], COUNT / SUM(I2) / SUM(R1) / MIN(I1) / MAX(I1), NCCL
gather of aggregates (always) and positions + tuples on rank 0.  4 G rows over 8 GPUs = 500 M rows (22 GB) per GPU;
the table is generated on the device per shard from the counter RNG and never exists on the host.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c5.py [--rows-per-gpu 500000000]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import mbcol
from mbcol import sharding
from bench import _CudaArray, measured_peak_gbs

N = mbcol._native
SEED = 20260101
DESCS = [(1, 4)] * 5 + [(2, 4)] * 2 + [(0, 16)]                  # I1..I5, R1, R2, S
AGGS = [(0, 0), (1, 1), (1, 5), (2, 0), (3, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=500_000_000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    ctx = mbcol.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rows = a.rows_per_gpu
    t = ctx.create_table(DESCS, rows, position_base=rank * rows)          # positions are int64 = shard base + local row
    for c in range(5):
        t.generate(c, 0, SEED, 1 << 20)
    t.generate(5, 1, SEED)
    t.generate(6, 1, SEED)
    t.generate(7, 2, SEED)
    r = float(np.sqrt(0.01))
    terms = [mbcol.Term(N.OP_LT, ("col", 0), ("int", int(np.ceil(r * (1 << 20)))), 0),
             mbcol.Term(N.OP_LT, ("col", 5), ("real", float(np.float32(r * 1000.0))), 1)]
    want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG

    def step():
        res = t.scan(terms, proj=[0, 1, 5], want=want, aggs=AGGS)
        ms = ctx.last_kernel_ms
        total = res.count
        if world > 1:
            mine = torch.as_tensor(_CudaArray(res.device_pointers()["aggs"], 9 * 8), device=dev).view(torch.int64)
            blocks = sharding.allgather_blocks(mine)
            cnts = [int(c) for c in blocks[:, 8].cpu()]
            total = sum(cnts)
            bufs = [(res.device_pointers()["positions"], 8)] + [res.column_device(i) for i in range(3)]
            locals_ = [(torch.as_tensor(_CudaArray(p, max(res.count, 1) * s), device=dev)[:res.count * s], s) for p, s in bufs]
            got = sharding.gather_rows_multi(locals_, cnts)      # 100 MB per rank: the grouped send/recv form wins here (profiles/README.md)
            if rank == 0:
                pos = got[0].view(torch.int64)
                assert pos.numel() == total and bool((pos[1:] > pos[:-1]).all())      # rank order = position order
        res.close()
        return ms, total

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kms = []
    for _ in range(a.steps):
        ms, total = step()
        kms.append(ms)
    e1.record()
    torch.cuda.synchronize()
    wall = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, _ = measured_peak_gbs()
        k = statistics.mean(kms)
        alg = rows * (12 + 0.01 * 20)                           # N/G * (12 + s * (8 + 12)) per GPU
        print(json.dumps({"config": "C5", "n_gpus": world, "rows_per_gpu": rows, "total_rows": rows * world, "selected": total,
                          "ms_per_step_wall": float(wall.item()) / a.steps, "scan_kernel_ms": k,
                          "rows_per_s_aggregate": rows * world * a.steps / (float(wall.item()) * 1e-3),
                          "per_gpu_achieved_gbs": alg / 1e9 / (k * 1e-3), "frac_of_measured_peak": alg / 1e9 / (k * 1e-3) / peak,
                          "gathered_bytes": total * 20}))
    t.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
