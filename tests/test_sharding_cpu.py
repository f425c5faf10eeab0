"""The multi-GPU host logic (TID-range shards, aggregate folding, ordered gather of positions/values on rank 0)
with world_size 2 and 3 over gloo on the CPU.  Each rank's local scan is done by the oracle here (the GPU
kernels are covered by the -m gpu tests); what is under test is mbcol.sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOTAL = 50_003
AGGS = [(0, 0), (1, 1), (1, 2), (2, 0), (3, 0)]


def _worker(rank, world, port, sel, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mbcol import sharding
        from oracle import oracle as orc
        from util import C2_DESCS, c2_columns, c2_terms
        lo, hi = sharding.shard_range(TOTAL, world, rank)
        cols = c2_columns(orc, hi - lo, position_base=lo)          # the shard regenerates its rows from the counter RNG
        res = orc.scan(C2_DESCS, cols, c2_terms(orc, sel), proj=[0, 1, 2, 3], aggs=AGGS)
        raw = []
        for (kind, col), (ai, af, av) in zip(AGGS, res["aggs"]):
            is_real = kind != 0 and C2_DESCS[col][0] == 2
            if is_real:
                raw.append(int(np.float64(af).view(np.int64)))
            else:
                raw.append(ai if av else (2**31 - 1 if kind == 2 else -2**31 if kind == 3 else 0))
        block = torch.tensor(raw + [res["count"]], dtype=torch.int64)
        blocks = sharding.allgather_blocks(block)
        folded, total = sharding.fold_aggregates(blocks, [k for k, _ in AGGS], [k != 0 and C2_DESCS[c][0] == 2 for k, c in AGGS])
        counts = [int(c) for c in blocks[:, -1]]
        pos = torch.from_numpy((res["positions"] + lo).astype(np.int64)).view(torch.uint8)
        allpos = sharding.gather_rows(pos, counts, 8)
        s_col = torch.from_numpy(np.ascontiguousarray(cols[3][res["positions"]])).reshape(-1)
        alls = sharding.gather_rows(s_col, counts, 16)
        packed = sharding.gather_rows_packed([(pos, 8), (s_col, 16)], counts)     # one message per rank: same concatenations
        dense = sharding.allgather_rows([(pos, 8), (s_col, 16)], counts)         # one all-gather of padded blocks
        if rank == 0:
            assert torch.equal(packed[0], allpos) and torch.equal(packed[1], alls)
            assert torch.equal(dense[0], allpos) and torch.equal(dense[1], alls)
            out.put({"folded": folded, "total": total, "positions": allpos.view(torch.int64).numpy().copy(),
                     "S": alls.numpy().reshape(-1, 16).copy()})
        else:
            assert packed is None and dense is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sel", [(2, 0.1), (3, 0.01)])
def test_sharded_scan_equals_single_table_scan(world, sel):
    from oracle import oracle as orc
    from util import C2_DESCS, c2_columns, c2_terms
    orc.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, sel, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cols = c2_columns(orc, TOTAL)
    exp = orc.scan(C2_DESCS, cols, c2_terms(orc, sel), proj=[0, 1, 2, 3], aggs=AGGS)
    assert got["total"] == exp["count"]
    np.testing.assert_array_equal(got["positions"], exp["positions"])          # rank order = position order
    np.testing.assert_array_equal(got["S"], cols[3][exp["positions"]])
    for (gv, gvalid), (ei, ef, ev), (kind, col) in zip(got["folded"], exp["aggs"], AGGS):
        assert gvalid == ev
        if kind != 0 and C2_DESCS[col][0] == 2:
            assert abs(gv - ef) <= 1e-6 * abs(ef)                                   # real SUM: association differs across shards
        else:
            assert gv == ei


def test_shard_ranges_cover_the_table():
    from mbcol import sharding
    for total in (0, 1, 8191, 8192, 50_003, 4_000_000_000):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo % sharding.TILE == 0 or lo == total for lo, _ in r)


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mbcol import sharding
        counts = [5, 0, 3][:world]                                  # rank 1 found nothing
        n = counts[rank]
        a = (torch.arange(n * 8, dtype=torch.int64) + 1000 * rank).to(torch.uint8)          # 8-byte rows
        b = (torch.arange(n * 20, dtype=torch.int64) * 3 + rank).to(torch.uint8)            # 20-byte rows
        forms = {"multi": sharding.gather_rows_multi([(a, 8), (b, 20)], counts),
                 "packed": sharding.gather_rows_packed([(a, 8), (b, 20)], counts),
                 "allgather": sharding.allgather_rows([(a, 8), (b, 20)], counts),
                 "everyone": sharding.allgather_rows([(a, 8), (b, 20)], counts, dst=None)}
        assert forms["everyone"] is not None and forms["everyone"][0].numel() == sum(counts) * 8
        if rank == 0:
            out.put({k: [t.numpy().copy() for t in v] for k, v in forms.items()})
        else:
            assert forms["multi"] is None and forms["packed"] is None and forms["allgather"] is None
    finally:
        dist.destroy_process_group()


def test_gather_forms_agree_with_an_empty_rank():
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, 29650, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    counts = [5, 0, 3]
    exp_a = np.concatenate([((np.arange(c * 8) + 1000 * r) % 256).astype(np.uint8) for r, c in enumerate(counts)])
    exp_b = np.concatenate([((np.arange(c * 20) * 3 + r) % 256).astype(np.uint8) for r, c in enumerate(counts)])
    for form, (a, b) in got.items():
        np.testing.assert_array_equal(a, exp_a, err_msg=form)
        np.testing.assert_array_equal(b, exp_b, err_msg=form)
