"""Shared helpers of the parity tests (host side only; no reference access)."""
import numpy as np

SEED = 20260101

# BASELINE config C2: I1, I2 int in [0, 2^20), R real in [0, 1000), S char(16)
C2_DESCS = [(1, 4), (1, 4), (2, 4), (0, 16)]
C2_NAMES = ["I1", "I2", "R", "S"]


def c2_columns(oracle, nrows, position_base=0, seed=SEED):
    return [oracle.synth_int(seed, 0, nrows, 1 << 20, position_base),
            oracle.synth_int(seed, 1, nrows, 1 << 20, position_base),
            oracle.synth_real(seed, 2, nrows, position_base),
            oracle.synth_str(seed, 3, nrows, 16, position_base)]


def c2_device_table(ctx, nrows, position_base=0, seed=SEED):
    t = ctx.create_table(C2_DESCS, nrows, position_base)
    t.generate(0, 0, seed, 1 << 20)
    t.generate(1, 0, seed, 1 << 20)
    t.generate(2, 1, seed)
    t.generate(3, 2, seed)
    return t


def c2_terms(oracle, selectivity):
    """{(I1,<,t1)}^{(R,<,t2)} with per-predicate selectivity sqrt(s)."""
    r = float(np.sqrt(selectivity))
    t1 = int(np.ceil(r * (1 << 20)))
    t2 = float(np.float32(r * 1000.0))
    return [oracle.Term(oracle.OP_LT, ("col", 0), ("int", t1), 0),
            oracle.Term(oracle.OP_LT, ("col", 2), ("real", t2), 1)]


C2_AGGS = [(0, 0), (1, 1), (1, 2), (2, 0), (3, 0)]      # COUNT, SUM(I2), SUM(R), MIN(I1), MAX(I1)


def load_table(ctx, descs, columns, position_base=0):
    from oracle import oracle as orc
    n = orc.nrows_of(descs, columns)
    t = ctx.create_table(descs, n, position_base)
    for c, col in enumerate(columns):
        t.load_column(c, col)
    return t


def check_result(oracle, res, exp, proj_descs, want_tuples=True, real_rel_tol=1e-6):
    """GPU Result vs oracle dict: bit-exact positions / values / tuple bytes, exact integer aggregates,
    real SUM within real_rel_tol relative (the tolerance BASELINE.json's north_star states)."""
    assert res.count == exp["count"]
    np.testing.assert_array_equal(res.positions(), exp["positions"])
    if proj_descs:
        offs = oracle.tuple_layout(proj_descs)
        for i, (t, w) in enumerate(proj_descs):
            got = res.column(i)
            field = exp["tuples"][:, offs[i]:offs[i + 1]] if exp["count"] else np.empty((0, offs[i + 1] - offs[i]), np.uint8)
            if t == 0:
                assert got.shape == (exp["count"], w)
                np.testing.assert_array_equal(got, field[:, 2:])
            else:
                np.testing.assert_array_equal(got.view(np.uint32), field.copy().view(">u4").reshape(-1).astype(np.uint32))
        if want_tuples:
            tup = res.tuples()
            if exp["count"]:
                np.testing.assert_array_equal(tup, exp["tuples"])
    for a, (ei, ef, ev) in enumerate(exp["aggs"]):
        gi, gf, gv = res.agg(a)
        assert gv == ev, f"agg {a} valid"
        if not ev:
            continue
        if float(ei) == ef:          # integral aggregate: exact
            assert gi == ei, f"agg {a}: {gi} != {ei}"
        else:
            assert abs(gf - ef) <= real_rel_tol * max(abs(ef), 1e-30), f"agg {a}: {gf} vs {ef}"


def check_sorted_like_golden(lines, golden_rows, nkeys, keys_lead_projection):
    """`sort` output against the reference's transcript.  The Java leaves the order of equal keys unspecified: the row
    MULTISET must match, and -- when the projected fields start with the sort keys -- the sequence of key tuples must be
    identical line by line."""
    assert sorted(lines) == sorted(golden_rows)
    if keys_lead_projection:
        key = lambda ln: ln.rsplit(" :", 1)[0].split(" ")[:nkeys]
        assert [key(a) for a in lines] == [key(b) for b in golden_rows]
