"""Edge cases and error behaviour of the C ABI (the Java shim maps non-zero statuses to the reference's checked
exceptions): limits, argument validation, PredEval's literal typing quirk, shapes that produce nothing."""
import numpy as np
import pytest

import mbcol
from mbcol import _native as N
from util import C2_DESCS, c2_columns, c2_terms, check_result, load_table

pytestmark = pytest.mark.gpu
ALL = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST


def test_argument_validation(ctx, oracle):
    t = load_table(ctx, [(1, 4), (0, 8)], [np.arange(10, dtype=np.int32), oracle.pack_strings(["a"] * 10, 8)])
    cases = [
        ([mbcol.Term(N.OP_EQ, ("col", 7), ("int", 1), 0)], [], N.ERR_ARG),                      # column out of range
        ([mbcol.Term(N.OP_EQ, ("col", 0), ("str", "x"), 0)], [], N.ERR_UNSUPPORTED),            # int column vs string literal
        ([mbcol.Term(N.OP_EQ, ("col", 1), ("int", 1), 0)], [], N.ERR_UNSUPPORTED),              # string column vs int literal
        ([mbcol.Term(N.OP_EQ, ("col", 1), ("str", "x" * 65), 0)], [], N.ERR_UNSUPPORTED),       # literal wider than 64 bytes
        ([mbcol.Term(11, ("col", 0), ("int", 1), 0)], [], N.ERR_ARG),                           # unknown operator
        ([mbcol.Term(N.OP_EQ, ("col", 0), ("int", 1), 0)], [5], N.ERR_ARG),                     # projection out of range
        ([mbcol.Term(N.OP_EQ, ("col", 0), ("int", i), i) for i in range(17)], [], N.ERR_UNSUPPORTED),   # > 16 terms
    ]
    for terms, proj, status in cases:
        with pytest.raises(mbcol.MbcError) as ei:
            t.scan(terms, proj=proj, want=ALL)
        assert ei.value.status == status, (terms[0], ei.value)
        assert ei.value.message
    with pytest.raises(mbcol.MbcError) as ei:
        t.scan([], want=N.WANT_AGG, aggs=[(N.AGG_SUM, 1)])                                       # SUM over a string column
    assert ei.value.status == N.ERR_UNSUPPORTED
    with pytest.raises(mbcol.MbcError):
        ctx.create_table([(2, 8)], 10)                                                           # real must be 4 bytes wide
    with pytest.raises(mbcol.MbcError):
        ctx.create_table([(0, 300)], 10)                                                         # char(300) unsupported
    t.close()


def test_literal_typing_follows_the_lhs(ctx, oracle):
    """PredEval compares under the type of the LEFT operand and reads the other side's 4 bytes with that type's
    getter (iterator/PredEval.java:60-123): an int column against a real literal compares the float's bit pattern."""
    n = 5000
    rng = np.random.default_rng(1)
    descs = [(1, 4), (2, 4)]
    cols = [rng.integers(0, 2_000_000_000, n).astype(np.int32), rng.random(n).astype(np.float32) * 10]
    t = load_table(ctx, descs, cols)
    for terms in ([oracle.Term(oracle.OP_LT, ("col", 0), ("real", 2.5), 0)],          # int col <  bits(2.5f) = 0x40200000
                  [oracle.Term(oracle.OP_GT, ("real", 5.0), ("col", 1), 0)],          # literal on the left: real compare
                  [oracle.Term(oracle.OP_LE, ("int", 1_000_000_000), ("col", 0), 0)],
                  [oracle.Term(oracle.OP_NOP, ("col", 0), ("int", 1), 0)],            # aopNOP / opRANGE never match
                  [oracle.Term(oracle.OP_RANGE, ("col", 0), ("int", 1), 0)],
                  [oracle.Term(oracle.OP_NOT, ("col", 0), ("int", int(cols[0][0])), 0)]):   # aopNOT behaves as !=
        exp = oracle.scan(descs, cols, terms, proj=[1, 0], aggs=[(0, 0), (1, 0), (1, 1)])
        res = t.scan(terms, proj=[1, 0], want=ALL, aggs=[(0, 0), (1, 0), (1, 1)])
        check_result(oracle, res, exp, [descs[1], descs[0]])
        res.close()
    t.close()


def test_shapes_that_return_nothing_or_everything(ctx, oracle):
    n = 12_345
    cols = c2_columns(oracle, n)
    t = load_table(ctx, C2_DESCS, cols)
    none = t.scan([oracle.Term(oracle.OP_LT, ("col", 0), ("int", -1), 0)], proj=[3, 0], want=ALL | N.WANT_BITMAP,
                  aggs=[(0, 0), (1, 1), (2, 0), (3, 2)])
    assert none.count == 0 and none.positions().size == 0 and none.tuples().shape[0] == 0 and not none.bitmap().any()
    assert none.agg(0) == (0, 0.0, True) and none.agg(1)[0] == 0 and none.agg(2)[2] is False and none.agg(3)[2] is False
    every = t.scan([oracle.Term(oracle.OP_GE, ("col", 0), ("int", 0), 0)], proj=[], want=N.WANT_POSITIONS | N.WANT_AGG | N.WANT_HOST,
                   aggs=[(0, 0)])
    assert every.count == n and every.agg(0)[0] == n
    np.testing.assert_array_equal(every.positions(), np.arange(n))
    only_aggs = t.scan([], want=N.WANT_AGG, aggs=[(1, 1), (3, 2)])                    # nothing materialised but aggregates
    assert only_aggs.agg(0)[0] == int(cols[1].astype(np.int64).sum()) and only_aggs.agg(1)[1] == float(cols[2].max())
    counts = t.scan(c2_terms(oracle, 0.1), want=N.WANT_AGG, aggs=[(0, 0), (0, 2)])      # COUNT only: no write pass at all
    exp = oracle.scan(C2_DESCS, cols, c2_terms(oracle, 0.1))["count"]
    assert counts.count == exp and counts.agg(0) == (exp, float(exp), True) and counts.agg(1)[0] == exp
    for r in (none, every, only_aggs, counts):
        r.close()
    t.close()


def test_many_conjuncts_and_wide_strings(ctx, oracle):
    """16 terms over 5 columns (more staged predicate columns than the TMA ring holds -> the rest read from HBM), and a
    char(40) column in both predicate and projection."""
    n = 70_000
    rng = np.random.default_rng(3)
    descs = [(1, 4)] * 5 + [(0, 40)]
    words = ["alpha", "alphabet", "beta", "a-very-long-string-value-of-39-bytes-..", "gamma"]
    cols = [rng.integers(0, 100, n).astype(np.int32) for _ in range(5)]
    cols.append(oracle.pack_strings([words[i] for i in rng.integers(0, len(words), n)], 40))
    t = load_table(ctx, descs, cols)
    terms = []
    for i in range(14):
        terms.append(oracle.Term(int(rng.integers(0, 6)), ("col", i % 5), ("int", int(rng.integers(20, 80))), i // 3))
    terms.append(oracle.Term(oracle.OP_GE, ("col", 5), ("str", "alphabet"), 5))
    terms.append(oracle.Term(oracle.OP_NE, ("col", 5), ("str", "gamma"), 5))
    exp = oracle.scan(descs, cols, terms, proj=[5, 4, 0], aggs=[(0, 0), (1, 3)])
    res = t.scan(terms, proj=[5, 4, 0], want=ALL, aggs=[(0, 0), (1, 3)])
    assert exp["count"] > 0
    check_result(oracle, res, exp, [descs[5], descs[4], descs[0]])
    res.close()
    t.close()


def test_deferred_results_match_synchronous(ctx, oracle):
    """Device-resident results of mbc_scan complete asynchronously: several scans are queued before the first count
    is read; counts, aggregates, device positions / values equal the synchronous (MBC_WANT_HOST) results."""
    import torch
    from bench import _CudaArray
    from util import C2_AGGS, c2_device_table, c2_terms
    n = 300_000
    t = c2_device_table(ctx, n)
    sels = (0.01, 0.1, 0.5, 0.1)
    want_dev = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG
    queued = [t.scan(c2_terms(oracle, s), proj=[1, 3], want=want_dev, aggs=C2_AGGS) for s in sels]
    for s, r in zip(sels, queued):
        ref = t.scan(c2_terms(oracle, s), proj=[1, 3], want=want_dev | N.WANT_HOST, aggs=C2_AGGS)
        assert r.count == ref.count > 0 and r.kernel_ms > 0
        for a in range(len(C2_AGGS)):
            assert r.agg(a) == ref.agg(a)
        ptrs = r.device_pointers()
        pos = torch.as_tensor(_CudaArray(ptrs["positions"], r.count * 8), device="cuda:0").view(torch.int64).cpu().numpy()
        np.testing.assert_array_equal(pos, ref.positions())
        p1, s1 = r.column_device(0)
        i2 = torch.as_tensor(_CudaArray(p1, r.count * s1), device="cuda:0").view(torch.int32).cpu().numpy()
        np.testing.assert_array_equal(i2, ref.column(0))
        p2, s2 = r.column_device(1)
        sv = torch.as_tensor(_CudaArray(p2, r.count * s2), device="cuda:0").cpu().numpy().reshape(-1, s2)
        np.testing.assert_array_equal(sv[:, :16], ref.column(1))
        ref.close()
    dropped = t.scan(c2_terms(oracle, 0.5), proj=[0], want=want_dev)      # freed without ever being read
    dropped.close()
    for r in queued:
        r.close()
    t.close()


def test_both_write_pass_forms(ctx, oracle, monkeypatch):
    """The write pass has a one-CTA-per-tile form and a persistent form (tables of >= 49152 tiles); both are forced
    here over the same mixed table: an all-qualifying stretch (dense groups), a sparse stretch and an empty one."""
    n = 400_000
    rng = np.random.default_rng(11)
    key = np.concatenate([np.zeros(150_000, np.int32), rng.integers(0, 100, 150_000).astype(np.int32),
                          np.full(100_000, 1000, np.int32)])
    descs = [(1, 4), (2, 4), (0, 16)]
    cols = [key, rng.random(n).astype(np.float32), oracle.synth_str(7, 3, n, 16, 0)]
    t = load_table(ctx, descs, cols)
    terms = [oracle.Term(oracle.OP_LT, ("col", 0), ("int", 3), 0)]
    aggs = [(0, 0), (1, 0), (1, 1), (2, 1), (3, 1)]
    exp = oracle.scan(descs, cols, terms, proj=[2, 1, 0], aggs=aggs)
    assert 150_000 < exp["count"] < 160_000
    for force in ("1", "1000000000"):
        monkeypatch.setenv("MBC_WRITE_PERSISTENT_TILES", force)
        # staged on: the fullest groups (persistent form: every group above the sparse limit) go through write_staged_kernel;
        # staged off: write_kernel gathers them tile by tile (the persistent form's static stride over the dense tiles)
        for staged in ("1", "0"):
            monkeypatch.setenv("MBC_WRITE_STAGED", staged)
            res = t.scan(terms, proj=[2, 1, 0], want=ALL, aggs=aggs)
            check_result(oracle, res, exp, [descs[2], descs[1], descs[0]])
            res.close()
    # a mid-density stretch (20 % of the middle rows) under both forms, staged on
    terms2 = [oracle.Term(oracle.OP_LT, ("col", 0), ("int", 20), 0)]
    exp2 = oracle.scan(descs, cols, terms2, proj=[0, 2], aggs=aggs)
    for force in ("1", "1000000000"):
        monkeypatch.setenv("MBC_WRITE_PERSISTENT_TILES", force)
        monkeypatch.setenv("MBC_WRITE_STAGED", "1")
        res = t.scan(terms2, proj=[0, 2], want=ALL, aggs=aggs)
        check_result(oracle, res, exp2, [descs[0], descs[2]])
        res.close()
    t.close()


def test_multibyte_utf8_strings_compare_like_string_compareto(ctx, oracle):
    """TupleUtils.java:79-81 compares Java Strings (UTF-16 code units); the columns hold the bytes Convert.setStrValue wrote
    (modified UTF-8).  For text without U+0000 the byte order of modified UTF-8 IS the UTF-16 code-unit order (2- and 3-byte
    sequences sort by code point; the BMP has no surrogates), so the kernel's byte compare over the zero-padded width equals
    compareTo -- checked against the oracle, which decodes to UTF-16 and compares like the Java."""
    rng = np.random.default_rng(11)
    words = ["abc", "abd", "ab", "é", "ébc", "eb", "z", "Zürich", "zürich", "日本", "日本語", "中", "ß", "straße", "strasse", "~", "߿", "ࠀ",
             "￮", "aé", "á", "€uro", "ñandú", "nandu"]
    n = 9001
    pick = rng.integers(0, len(words), n)
    col = oracle.pack_strings([words[i] for i in pick], 16)
    other = oracle.pack_strings([words[i] for i in rng.integers(0, len(words), n)], 16)
    ints = rng.integers(0, 5, n).astype(np.int32)
    descs = [(0, 16), (0, 16), (1, 4)]
    t = load_table(ctx, descs, [col, other, ints])
    want = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_AGG | N.WANT_HOST
    for lit in ("é", "z", "日本", "straße", "á", "ࠀ", "abc"):
        for op in range(7):
            terms = [oracle.Term(op, ("col", 0), ("str", lit), 0)]
            exp = oracle.scan(descs, [col, other, ints], terms, proj=[0, 2], aggs=[(0, 0)])
            res = t.scan(terms, proj=[0, 2], want=want, aggs=[(0, 0)])
            check_result(oracle, res, exp, [descs[0], descs[2]])
            res.close()
    for op in (0, 1, 5):                                        # column against column, literal on the left
        terms = [oracle.Term(op, ("col", 0), ("col", 1), 0), oracle.Term(op, ("str", "ñandú"), ("col", 1), 1)]
        exp = oracle.scan(descs, [col, other, ints], terms, proj=[1, 0])
        res = t.scan(terms, proj=[1, 0], want=N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_TUPLES | N.WANT_HOST)
        check_result(oracle, res, exp, [descs[1], descs[0]])
        res.close()
    t.close()
