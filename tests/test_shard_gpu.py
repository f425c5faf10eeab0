"""mbc_shard_*: TID-range shards gathered into the root's window through peer-memory stores, against the oracle.

One process drives every rank here (the shape of a single JVM with several GPUs): one Context per rank, on distinct
GPUs when the box has them, else all on cuda:0 (the window is then ordinary device memory and the pushes run one after the
other -- the kernels, offsets, flags and folds are the same).  The one-process-per-GPU form (CUDA IPC handles) is
exercised by bench.py --gpus N.  Needs a B200."""
import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import C2_AGGS, C2_DESCS, c2_columns, c2_terms

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_scan_gathers_the_whole_tables_result(oracle, world):
    nrows_rank = 250_000 + 4096 * 3
    total = nrows_rank * world
    cols = c2_columns(oracle, total)
    ndev = _ndev()
    ctxs = [mbcol.Context(r % ndev) for r in range(world)]
    tables = []
    for r, c in enumerate(ctxs):
        t = c.create_table(C2_DESCS, nrows_rank, position_base=r * nrows_rank)
        for k in range(4):
            t.generate(k, (0, 0, 1, 2)[k], 20260101, 1 << 20 if k < 2 else 0)
        tables.append(t)
    proj = [0, 1, 2, 3]
    pdescs = [C2_DESCS[c] for c in proj]
    shards = [mbcol.Shard(c, r, world) for r, c in enumerate(ctxs)]
    shards[0].create_window(int(0.6 * total) + 4096, pdescs)
    for s in shards[1:]:
        s.attach(shards[0])
    want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG
    for step, sel in enumerate((0.01, 0.5, 0.1, 0.01, 0.5)):       # five steps: both window slots are reused, releases are needed
        terms = c2_terms(oracle, sel)
        exp = oracle.scan(C2_DESCS, cols, terms, proj=proj, aggs=C2_AGGS, nthreads=oracle.max_threads())
        results = [t.scan(terms, proj=proj, want=want, aggs=C2_AGGS) for t in tables]
        for r in range(world):                                      # rank order; every push has finished before the next starts
            shards[r].gather(results[r], beside_next_scan=bool(step & 1), copy_engines=step in (2, 3))
            shards[r].fence()
            ctxs[r].sync()
        tot, counts = shards[0].collect()
        assert tot == exp["count"] and counts == [res.count for res in results]
        pos = shards[0].positions(0, tot)
        np.testing.assert_array_equal(pos, exp["positions"])
        for i, c in enumerate(proj):
            got = shards[0].column(i, 0, tot)
            if C2_DESCS[c][0] == 2:
                np.testing.assert_array_equal(got.view(np.uint32), cols[c][exp["positions"]].view(np.uint32))
            else:
                np.testing.assert_array_equal(got, cols[c][exp["positions"]])
        for a, (kind, col) in enumerate(C2_AGGS):
            ctype = N.ATTR_INTEGER if kind == 0 else C2_DESCS[col][0]
            gi, gf, gv = shards[0].agg(a, kind, ctype)
            ei, ef, ev = exp["aggs"][a]
            assert gv == ev
            if ev:
                assert gi == ei if float(ei) == ef else abs(gf - ef) <= 1e-6 * max(abs(ef), 1e-30)
        shards[0].release()
        for res in results:
            res.close()
    for s in shards[1:] + shards[:1]:
        s.close()
    for t in tables:
        t.close()
    for c in ctxs:
        c.close()


def test_window_overflow_is_reported(oracle):
    ctx = mbcol.Context(0)
    t = ctx.create_table(C2_DESCS, 100_000)
    for k in range(4):
        t.generate(k, (0, 0, 1, 2)[k], 20260101, 1 << 20 if k < 2 else 0)
    sh = mbcol.Shard(ctx, 0, 1)
    sh.create_window(1000, [C2_DESCS[0]])
    res = t.scan([], proj=[0], want=N.WANT_POSITIONS | N.WANT_COLUMNS)
    sh.gather(res)
    with pytest.raises(mbcol.MbcError) as e:
        sh.collect()
    assert e.value.status == N.ERR_UNSUPPORTED and "did not fit" in e.value.message
    np.testing.assert_array_equal(sh.positions(0, 1000), np.arange(1000))     # what fitted is intact
    res.close(); sh.close(); t.close(); ctx.close()
