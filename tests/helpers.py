"""Shared helpers of the parity tests: synthetic measure data through the ORACLE's part writer (the
restated reference writer), the same query through oracle and GPU, and the comparison rules of the
parity contract (bit-exact int64 and float64 min/max, <= 1e-9 relative for float64 sum/mean)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

T0 = 1_700_000_000_000_000_000
STEP = 60_000_000_000


def build_part(sids, ts, ver, fields=(), families=()):
    b = O.PartBuilder()
    b.append(sids, ts, ver, list(fields), list(families))
    return b.finish()


def grid(n_series, n_pts, sid0=1, sid_step=1, t0=T0, step=STEP):
    sids = np.repeat(sid0 + np.arange(n_series, dtype=np.uint64) * sid_step, n_pts)
    ts = np.tile(t0 + np.arange(n_pts, dtype=np.int64) * step, n_series)
    return sids, ts, np.ones(sids.size, dtype=np.int64)


def to_gpu_query(bydb, handles, oq: O.Query):
    return bydb.Query(parts=handles, series_ids=np.asarray(oq.sids, dtype=np.uint64), aggs=list(oq.aggs),
                      series_group=None if oq.groups is None else np.asarray(oq.groups, dtype=np.int32),
                      n_groups=oq.n_groups, tmin=oq.tmin, tmax=oq.tmax,
                      preds=[bydb.Pred(p.family, p.tag, p.op, p.value) for p in oq.preds],
                      top_n=oq.top_n, top_agg=oq.top_agg, top_desc=oq.top_desc)


def assert_parity(got, want, aggs, ctx=""):
    assert got.group_id.tolist() == want.group_id.tolist(), f"{ctx}: group ids {got.group_id} vs {want.group_id}"
    assert got.rows.tolist() == want.rows.tolist(), f"{ctx}: rows"
    assert got.is_float.tolist() == want.is_float.tolist(), f"{ctx}: output typing"
    for a, (fname, func) in enumerate(aggs):
        if not want.is_float[a]:
            assert got.val_i64[:, a].tolist() == want.val_i64[:, a].tolist(), f"{ctx}: int64 agg {a} ({fname},{func}) must be bit-exact"
        elif func in (O.AGG_MIN, O.AGG_MAX):
            assert got.val_f64[:, a].view(np.uint64).tolist() == want.val_f64[:, a].view(np.uint64).tolist(), \
                f"{ctx}: float64 min/max agg {a} must be bit-exact: {got.val_f64[:, a]} vs {want.val_f64[:, a]}"
        else:
            g, w = got.val_f64[:, a], want.val_f64[:, a]
            tol = 1e-9 * np.maximum(np.abs(w), 1e-300)
            assert (np.abs(g - w) <= tol).all(), f"{ctx}: float64 sum/mean agg {a} beyond 1e-9 relative: {g} vs {w}"


def run_both(bydb, ctx, parts, oq: O.Query, part_id0=1000):
    """parts: list of oracle Part objects. Returns (gpu_result, oracle_result)."""
    handles = [ctx.register_part(part_id0 + i, p.files()) for i, p in enumerate(parts)]
    try:
        got = ctx.scan_agg(to_gpu_query(bydb, handles, oq))
    finally:
        for h in handles:
            ctx.release_part(h)
    want = O.run_query(oq)
    return got, want


# ------------------------------------------------------------------ the reference's end-to-end cases (tests/golden/e2e_cases.json)
import json as _json
import os as _os

_E2E = _json.load(open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "e2e_cases.json")))
E2E_CASES = sorted(_E2E)
_AGG = {"SUM": O.AGG_SUM, "COUNT": O.AGG_COUNT, "MIN": O.AGG_MIN, "MAX": O.AGG_MAX, "MEAN": O.AGG_MEAN}
_OPS = {"BINARY_OP_EQ": O.OP_EQ, "BINARY_OP_NE": O.OP_NE, "BINARY_OP_LT": O.OP_LT, "BINARY_OP_LE": O.OP_LE, "BINARY_OP_GT": O.OP_GT,
        "BINARY_OP_GE": O.OP_GE}


def load_e2e_case(name):
    """-> (oracle part, oracle query, group names by dense id, want rows, ordered?) for one case of the reference's
    test/cases/measure suite.  Series = distinct entity tuples (ids in order of first appearance), timestamps one
    interval apart in data order (test/cases/measure/data/data.go:238-252), every tag also stored as a tag column."""
    c = _E2E[name]
    tags, ent = c["tags"], c["entity"]
    series = {}
    for r in c["rows"]:
        series.setdefault(tuple(r["tags"][tags.index(t)] for t in ent), len(series) + 1)
    n = len(c["rows"])
    sid = np.array([series[tuple(r["tags"][tags.index(t)] for t in ent)] for r in c["rows"]], dtype=np.uint64)
    ts = T0 + np.arange(n, dtype=np.int64) * STEP
    order = np.lexsort((ts, sid))
    rows = [c["rows"][i] for i in order]
    fields = []
    for fi, f in enumerate(c["fields"]):
        vals = [r["fields"][fi] for r in rows]
        fields.append((f["name"], O.VT_FLOAT64 if f["type"] == "float" else O.VT_INT64,
                       np.array(vals, dtype=np.float64 if f["type"] == "float" else np.int64), None))
    tag_cols = [(t, O.VT_STR, [r["tags"][ti].encode() for r in rows], None) for ti, t in enumerate(tags)]
    part = build_part(sid[order], ts[order], np.ones(n, np.int64), fields, [(c["family"], tag_cols)])
    q = c["query"]
    gi = tags.index(q["group_by"])
    usid = np.unique(sid)
    names, gid_of_series = [], []
    for s in usid.tolist():
        vals = {r["tags"][gi] for r, rs in zip(c["rows"], sid.tolist()) if rs == s}
        assert len(vals) == 1, "group-by tag must be constant per series for the series->group table"
        v = vals.pop()
        if v not in names:
            names.append(v)
        gid_of_series.append(names.index(v))
    preds = []
    if q["criteria"]:
        preds.append(O.Pred(c["family"], q["criteria"]["tag"], _OPS[q["criteria"]["op"]], q["criteria"]["value"].encode()))
    oq = O.Query([part], usid, [(q["field"], _AGG[q["agg"]])], groups=np.array(gid_of_series, dtype=np.int32), n_groups=len(names), preds=preds,
                 top_n=q["top"]["n"] if q["top"] else 0, top_agg=0, top_desc=q["top"]["desc"] if q["top"] else True)
    return part, oq, names, c["want"], bool(q["top"])


def check_e2e_rows(res, names, want, ordered, ctx=""):
    got = [(names[g], float(res.val_f64[i, 0]) if res.is_float[0] else int(res.val_i64[i, 0])) for i, g in enumerate(res.group_id.tolist())]
    exp = [(w["group"], w["value"]) for w in want]
    if not ordered:
        got, exp = sorted(got), sorted(exp)
    assert [g for g, _ in got] == [g for g, _ in exp], f"{ctx}: groups {got} vs {exp}"
    for (g, a), (_, b) in zip(got, exp):
        # the row path's COUNT of a float field is a float; the vectorized path this library mirrors returns int64
        assert abs(float(a) - float(b)) <= 1e-9 * max(abs(float(b)), 1e-300), f"{ctx}: {g}: {a} vs {b}"


# ------------------------------------------------------------------ the fixtures of banyand/measure/query_test.go (tstable_test.go:333-487)
def query_test_fixture(which):
    """dpsTS1 / dpsTS11 / dpsTS2 reduced to what the hot path reads: series 1..3 at one timestamp with versions, an int64 and a
    float64 field (series 2 carries no field, series 3 only the int field -> null cells), and the int64 tag."""
    spec = {
        "dpsTS1": dict(ts=1, versions=[1, 2, 3], intField=[1110, None, 1110], floatField=[1221233.343, None, None], intTag=[10, None, None]),
        "dpsTS11": dict(ts=1, versions=[0, 1, 2], intField=[3330, None, 4440], floatField=[3663699.029, None, None], intTag=[30, None, None]),
        "dpsTS2": dict(ts=2, versions=[4, 5, 6], intField=[3330, None, 4440], floatField=[3663699.029, None, None], intTag=[30, None, None]),
    }[which]
    sids = np.array([1, 2, 3], dtype=np.uint64)
    ts = np.full(3, spec["ts"], dtype=np.int64)
    ver = np.array(spec["versions"], dtype=np.int64)

    def col(vals, dtype):
        nulls = np.array([v is None for v in vals], dtype=np.uint8)
        return np.array([0 if v is None else v for v in vals], dtype=dtype), nulls
    iv, inull = col(spec["intField"], np.int64)
    fv, fnull = col(spec["floatField"], np.float64)
    tv, tnull = col(spec["intTag"], np.int64)
    return build_part(sids, ts, ver, [("intField", O.VT_INT64, iv, inull), ("floatField", O.VT_FLOAT64, fv, fnull)],
                      [("singleTag", [("intTag", O.VT_INT64, tv, tnull)])])


# expected per-series aggregates of the cases of TestQueryResult (query_test.go:48-1364), derived from the rows its `want` keeps:
# (parts, {sid: (rows, sum(intField), count(intField), max(floatField) or None)}, kept versions per sid)
QUERY_TEST_CASES = {
    "duplicated data": (["dpsTS1", "dpsTS1"], {1: (1, 1110, 1, 1221233.343), 2: (1, 0, 0, None), 3: (1, 1110, 1, None)}, {1: [1], 2: [2], 3: [3]}),
    "different version 1": (["dpsTS1", "dpsTS11"], {1: (1, 1110, 1, 1221233.343), 2: (1, 0, 0, None), 3: (1, 1110, 1, None)}, {1: [1], 2: [2], 3: [3]}),
    "different version 2": (["dpsTS11", "dpsTS1"], {1: (1, 1110, 1, 1221233.343), 2: (1, 0, 0, None), 3: (1, 1110, 1, None)}, {1: [1], 2: [2], 3: [3]}),
    "multiple data 1": (["dpsTS1", "dpsTS2"], {1: (2, 4440, 2, 3663699.029), 2: (2, 0, 0, None), 3: (2, 5550, 2, None)}, {1: [1, 4], 2: [2, 5], 3: [3, 6]}),
    "multiple data 2": (["dpsTS2", "dpsTS1"], {1: (2, 4440, 2, 3663699.029), 2: (2, 0, 0, None), 3: (2, 5550, 2, None)}, {1: [1, 4], 2: [2, 5], 3: [3, 6]}),
}
QUERY_TEST_AGGS = [("intField", O.AGG_SUM), ("intField", O.AGG_COUNT), ("floatField", O.AGG_MAX), ("floatField", O.AGG_COUNT)]


def check_query_test_case(res, expect, ctx=""):
    assert res.group_id.tolist() == [0, 1, 2], ctx
    for g, sid in enumerate((1, 2, 3)):
        rows, isum, icnt, fmax = expect[sid]
        assert int(res.rows[g]) == rows, f"{ctx}: rows of series {sid}"
        assert int(res.val_i64[g, 0]) == isum and int(res.val_i64[g, 1]) == icnt, f"{ctx}: intField of series {sid}"
        assert int(res.val_i64[g, 3]) == (0 if fmax is None else rows), f"{ctx}: count(floatField) of series {sid}"
        if fmax is not None:
            assert float(res.val_f64[g, 2]) == fmax, f"{ctx}: max(floatField) of series {sid}"


# ------------------------------------------------------------------ banyand/measure/part_iter_test.go Test_partIter_nextBlock
def part_iter_fixture():
    """`dps` of part_test.go:130-133 reduced to what block selection reads: series 1,1,2,2,3,3 at timestamps 1,2,8,10,100,220."""
    sids = np.array([1, 1, 2, 2, 3, 3], dtype=np.uint64)
    ts = np.array([1, 2, 8, 10, 100, 220], dtype=np.int64)
    return build_part(sids, ts, np.arange(1, 7, dtype=np.int64), [("intField", O.VT_INT64, np.arange(6) * 10, None)])


# (query series, expected series of the selected blocks) over [1, 220] -- every block holds 2 rows (part_iter_test.go:43-110)
PART_ITER_CASES = [([1, 2, 3], [1, 2, 3]), ([], []), ([1], [1]), ([1, 3], [1, 3]), ([4], []), ([4, 5, 6], []), ([1, 4], [1])]
