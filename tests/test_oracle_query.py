"""Oracle query path (block selection, time trim, version dedup, row predicates, group-by, the five
aggregation functions, Top) checked against an independent numpy model and against the known answers
in the reference's pkg/query/vectorized/measure/aggregation_test.go and
pkg/query/aggregation/function.go semantics.  No GPU needed."""
import numpy as np
import pytest

from oracle import oracle as O

T0 = 1_700_000_000_000_000_000
STEP = 60_000_000_000


def _build(sids, ts, ver, f_int=None, f_flt=None, region=None, code=None, nulls=None):
    b = O.PartBuilder()
    fields = []
    if f_flt is not None:
        fields.append(("latency", O.VT_FLOAT64, f_flt, None))
    if f_int is not None:
        fields.append(("calls", O.VT_INT64, f_int, nulls))
    fams = []
    cols = []
    if region is not None:
        cols.append(("region", O.VT_STR, region, None))
    if code is not None:
        cols.append(("code", O.VT_INT64, code, None))
    if cols:
        fams.append(("default", cols))
    b.append(sids, ts, ver, fields, fams)
    return b.finish()


def _synthetic(n_series=7, n_pts=300, seed=1, with_null=False):
    rng = np.random.default_rng(seed)
    sids = np.repeat(np.arange(1, n_series + 1, dtype=np.uint64) * 11, n_pts)
    ts = np.tile(T0 + np.arange(n_pts, dtype=np.int64) * STEP, n_series)
    ver = np.ones(sids.size, dtype=np.int64)
    f_flt = np.round(25 + rng.normal(0, 5, sids.size), 2)
    f_int = rng.integers(-1000, 1000, sids.size)
    region = [b"r%d" % (i % 8) for i in rng.integers(0, 8, sids.size)]
    code = rng.integers(0, 5, sids.size) * 100
    nulls = (rng.random(sids.size) < 0.1) if with_null else None
    return sids, ts, ver, f_int, f_flt, region, code, nulls


def test_known_answers_aggregation_test_go():
    # aggregation_test.go:82-120: sum a=3, b=7 ; :150-200: count a=3, b=1
    sids = np.array([1, 2, 1, 2], dtype=np.uint64)
    ts = T0 + np.array([0, 0, 1, 1], dtype=np.int64) * STEP
    part = _build(sids, ts, np.ones(4, np.int64), f_int=np.array([1, 3, 2, 4]))
    r = O.run_query(O.Query([part], [1, 2], [("calls", O.AGG_SUM), ("calls", O.AGG_COUNT), ("calls", O.AGG_MIN),
                                            ("calls", O.AGG_MAX), ("calls", O.AGG_MEAN)],
                            groups=[0, 1], n_groups=2))
    assert r.group_id.tolist() == [0, 1]
    assert r.val_i64[0].tolist() == [3, 2, 1, 2, 1]   # mean = 3/2 -> 1 (integer division)
    assert r.val_i64[1].tolist() == [7, 2, 3, 4, 3]   # mean = 7/2 -> 3
    assert not r.is_float.any()


def test_mean_quirks_function_go():
    # function.go:31-40: result < 1 is clamped to 1; float mean
    sids = np.array([1, 1, 2, 2], dtype=np.uint64)
    ts = T0 + np.array([0, 1, 0, 1], dtype=np.int64) * STEP
    part = _build(sids, ts, np.ones(4, np.int64), f_int=np.array([-5, 2, 10, 11]), f_flt=np.array([0.25, 0.5, 1.5, 4.0]))
    r = O.run_query(O.Query([part], [1, 2], [("calls", O.AGG_MEAN), ("latency", O.AGG_MEAN), ("latency", O.AGG_COUNT)],
                            groups=[0, 1], n_groups=2))
    assert r.val_i64[:, 0].tolist() == [1, 10]          # (-3/2 -> -1) < 1 -> 1 ; 21/2 -> 10
    assert r.val_f64[:, 1].tolist() == [1.0, 2.75]      # 0.375 < 1 -> 1
    assert r.is_float.tolist() == [False, True, False]  # COUNT is always int64 (aggregation.go:425-430)
    assert r.val_i64[:, 2].tolist() == [2, 2]


def test_query_matches_numpy_model():
    sids, ts, ver, f_int, f_flt, region, code, _ = _synthetic()
    part = _build(sids, ts, ver, f_int, f_flt, region, code)
    assert part.meta()["total_count"] == sids.size
    usid = np.unique(sids)
    groups = (np.arange(usid.size) % 3).astype(np.int32)
    tmin, tmax = T0 + 40 * STEP, T0 + 220 * STEP
    q = O.Query([part], usid, [("latency", O.AGG_SUM), ("latency", O.AGG_MAX), ("latency", O.AGG_MIN),
                               ("calls", O.AGG_SUM), ("calls", O.AGG_MIN), ("calls", O.AGG_MAX), ("latency", O.AGG_MEAN)],
                groups=groups, n_groups=3, tmin=tmin, tmax=tmax,
                preds=[O.Pred("default", "region", O.OP_EQ, b"r3"), O.Pred("default", "code", O.OP_GE, 200)])
    r = O.run_query(q)
    reg = np.array(region)
    m = (ts >= tmin) & (ts <= tmax) & (reg == b"r3") & (code >= 200)
    gid_of = {int(s): int(g) for s, g in zip(usid, groups)}
    g = np.array([gid_of[int(s)] for s in sids])
    for row, grp in enumerate(r.group_id.tolist()):
        mm = m & (g == grp)
        assert r.rows[row] == mm.sum()
        assert r.val_f64[row, 0] == pytest.approx(f_flt[mm].sum(), rel=1e-12)
        assert r.val_f64[row, 1] == f_flt[mm].max()
        assert r.val_f64[row, 2] == f_flt[mm].min()
        assert r.val_i64[row, 3] == f_int[mm].sum()
        assert r.val_i64[row, 4] == f_int[mm].min()
        assert r.val_i64[row, 5] == f_int[mm].max()
        mean = f_flt[mm].sum() / mm.sum()
        assert r.val_f64[row, 6] == pytest.approx(max(mean, 1.0), rel=1e-12)
    assert r.rows_matched == m.sum()


def test_sequential_float_sum_is_row_order():
    # function.go:133-135: s.sum += val in scan order (series-major, then time)
    sids, ts, ver, _, f_flt, _, _, _ = _synthetic(n_series=3, n_pts=500, seed=3)
    part = _build(sids, ts, ver, f_flt=f_flt)
    r = O.run_query(O.Query([part], np.unique(sids), [("latency", O.AGG_SUM)]))
    acc = 0.0
    for v in f_flt.tolist():   # rows were generated already sorted (sid, ts)
        acc += v
    assert r.val_f64[0, 0] == acc


def test_version_dedup_across_parts():
    # query.go:995-1004 / query_batch.go:151-161: duplicate (sid, ts) keeps the highest version
    n = 50
    ts = T0 + np.arange(n, dtype=np.int64) * STEP
    sid = np.full(n, 7, dtype=np.uint64)
    p1 = _build(sid, ts, np.full(n, 1, np.int64), f_int=np.arange(n))
    # second part overwrites the even timestamps with version 2 and adds a stale version-0 copy of the odd ones
    ts2 = np.concatenate([ts[::2], ts[1::2]])
    ver2 = np.concatenate([np.full(n // 2, 2, np.int64), np.zeros(n // 2, np.int64)])
    val2 = np.concatenate([np.arange(n)[::2] + 1000, np.arange(n)[1::2] - 5000])
    p2 = _build(np.full(n, 7, np.uint64), ts2, ver2, f_int=val2)
    r = O.run_query(O.Query([p1, p2], [7], [("calls", O.AGG_SUM), ("calls", O.AGG_COUNT)]))
    want = (np.arange(n)[::2] + 1000).sum() + np.arange(n)[1::2].sum()
    assert r.val_i64[0].tolist() == [want, n]
    rows = O.scan_rows(O.Query([p1, p2], [7], [("calls", O.AGG_SUM)]))
    assert rows["ts"].tolist() == ts.tolist()
    assert rows["version"].tolist() == [2 if i % 2 == 0 else 1 for i in range(n)]


def test_in_part_duplicates_dropped_at_write():
    # part.go:176-190: same (sid, ts) inside one flush keeps the first after sort = highest version
    ts = T0 + np.array([0, 1, 1, 2], dtype=np.int64) * STEP
    part = _build(np.full(4, 3, np.uint64), ts, np.array([1, 1, 9, 1], np.int64), f_int=np.array([10, 20, 30, 40]))
    assert part.meta()["total_count"] == 3
    rows = O.scan_rows(O.Query([part], [3], [("calls", O.AGG_SUM)]))
    assert rows["fields"][0][2].tolist() == [10, 30, 40]


def test_null_cells_are_skipped_but_rows_counted():
    sids, ts, ver, f_int, f_flt, _, _, nulls = _synthetic(n_series=2, n_pts=100, seed=5, with_null=True)
    part = _build(sids, ts, ver, f_int=f_int, f_flt=f_flt, nulls=nulls)
    r = O.run_query(O.Query([part], np.unique(sids), [("calls", O.AGG_SUM), ("calls", O.AGG_COUNT), ("latency", O.AGG_COUNT)]))
    assert r.val_i64[0].tolist() == [f_int[~nulls].sum(), (~nulls).sum(), sids.size]
    assert r.rows[0] == sids.size


def test_block_cut_and_time_prune():
    # measure.go:41-46 / part.go:192-199: a block is cut when rows > 8192 -> 8193-row first block
    n = 20000
    ts = T0 + np.arange(n, dtype=np.int64) * STEP
    part = _build(np.full(n, 5, np.uint64), ts, np.ones(n, np.int64), f_int=np.arange(n))
    assert part.meta()["blocks_count"] == 3
    r = O.run_query(O.Query([part], [5], [("calls", O.AGG_SUM)], tmin=int(ts[8193]), tmax=int(ts[8200])))
    assert r.blocks_scanned == 1 and r.rows_scanned == 8193
    assert r.val_i64[0, 0] == np.arange(8193, 8201).sum()
    r = O.run_query(O.Query([part], [5], [("calls", O.AGG_COUNT)], tmin=int(ts[-1]) + 1))
    assert r.group_id.size == 0


def test_top_n_desc_and_ties():
    # top.go:145-214: largest N first; ties -> earlier row (lower group) wins
    sids = np.arange(1, 9, dtype=np.uint64)
    vals = np.array([5, 9, 9, 1, 7, 9, 3, 7])
    part = _build(sids, np.full(8, T0, np.int64), np.ones(8, np.int64), f_int=vals)
    r = O.run_query(O.Query([part], sids, [("calls", O.AGG_SUM)], groups=np.arange(8, dtype=np.int32), n_groups=8,
                            top_n=4, top_desc=True))
    assert r.group_id.tolist() == [1, 2, 5, 4]
    r = O.run_query(O.Query([part], sids, [("calls", O.AGG_SUM)], groups=np.arange(8, dtype=np.int32), n_groups=8,
                            top_n=2, top_desc=False))
    assert r.group_id.tolist() == [3, 6]


def test_threaded_modes_agree():
    sids, ts, ver, f_int, f_flt, region, code, _ = _synthetic(n_series=40, n_pts=400, seed=9)
    part = _build(sids, ts, ver, f_int, f_flt, region, code)
    usid = np.unique(sids)
    groups = (np.arange(usid.size) % 5).astype(np.int32)
    base = dict(groups=groups, n_groups=5, preds=[O.Pred("default", "region", O.OP_NE, b"r1")])
    aggs = [("latency", O.AGG_MAX), ("calls", O.AGG_SUM), ("latency", O.AGG_MEAN)]
    a = O.run_query(O.Query([part], usid, aggs, threads=1, **base))
    b = O.run_query(O.Query([part], usid, aggs, threads=4, **base))
    c = O.run_query(O.Query([part], usid, aggs, threads=4, per_thread_partials=True, **base))
    assert (a.val_i64 == b.val_i64).all() and (a.val_f64 == b.val_f64).all()
    assert (a.val_i64 == c.val_i64).all()
    np.testing.assert_allclose(a.val_f64, c.val_f64, rtol=1e-12)


def test_part_reopen_from_files():
    sids, ts, ver, f_int, f_flt, region, code, _ = _synthetic(n_series=4, n_pts=50, seed=2)
    part = _build(sids, ts, ver, f_int, f_flt, region, code)
    files = part.files()
    assert set(files) == {"meta.bin", "primary.bin", "timestamps.bin", "fv.bin", "default.tf", "default.tfm"}
    again = O.Part.open(files)
    q = lambda p: O.run_query(O.Query([p], np.unique(sids), [("latency", O.AGG_SUM), ("calls", O.AGG_MAX)]))
    assert q(part).val_f64.tolist() == q(again).val_f64.tolist()
    assert again.meta()["total_count"] == sids.size


@pytest.mark.parametrize("case", sorted(__import__("tests.helpers", fromlist=["x"]).QUERY_TEST_CASES))
def test_reference_query_test_fixtures(case):
    # banyand/measure/query_test.go TestQueryResult on dpsTS1 / dpsTS11 / dpsTS2 (tstable_test.go:333-487): which rows survive the
    # cross-part merge (highest version per (series, timestamp), either part order) and what they aggregate to
    from tests.helpers import QUERY_TEST_AGGS, QUERY_TEST_CASES, check_query_test_case, query_test_fixture
    names, expect, versions = QUERY_TEST_CASES[case]
    parts = [query_test_fixture(n) for n in names]
    q = O.Query(parts, [1, 2, 3], QUERY_TEST_AGGS, groups=np.arange(3, dtype=np.int32), n_groups=3, tmin=1, tmax=2)
    check_query_test_case(O.run_query(q), expect, case)
    rows = O.scan_rows(q)
    for sid, vs in versions.items():
        assert rows["version"][rows["sid"] == sid].tolist() == vs, f"{case}: versions kept for series {sid}"


@pytest.mark.parametrize("vals,n,asc,want", [
    ([5, 2, 8, 1, 7, 3], 3, True, [1, 2, 3]),      # top_test.go:71-88  TestBatchTop_AscendingHeap_KeepsLowestN
    ([5, 2, 8, 1, 7, 3], 3, False, [8, 7, 5]),     # top_test.go:90-104 TestBatchTop_DescendingHeap_KeepsHighestN
    ([3, 1, 2], 5, True, [1, 2, 3]),               # top_test.go:106-123 TestBatchTop_FewerInputThanN_ReturnsAll_InOrder
])
def test_top_known_answers_top_test_go(vals, n, asc, want):
    sids = np.arange(1, len(vals) + 1, dtype=np.uint64)
    part = _build(sids, np.full(len(vals), T0, np.int64), np.ones(len(vals), np.int64), f_int=np.array(vals))
    r = O.run_query(O.Query([part], sids, [("calls", O.AGG_SUM)], groups=np.arange(len(vals), dtype=np.int32), n_groups=len(vals),
                            top_n=n, top_desc=not asc))
    assert r.val_i64[:, 0].tolist() == want


def test_top_tie_breaker_stable_top_test_go():
    # top_test.go:125-158 TestBatchTop_TieBreaker_Stable: all values equal, only the first two rows are retained
    sids = np.arange(1, 5, dtype=np.uint64)
    part = _build(sids, np.full(4, T0, np.int64), np.ones(4, np.int64), f_int=np.full(4, 5))
    for desc in (False, True):
        r = O.run_query(O.Query([part], sids, [("calls", O.AGG_SUM)], groups=np.arange(4, dtype=np.int32), n_groups=4, top_n=2, top_desc=desc))
        assert r.group_id.tolist() == [0, 1]


def test_equal_version_duplicates_keep_the_earlier_part():
    # the reference leaves equal (series, timestamp, version) rows of different parts to the heap's internal order
    # (query.go:912-942, 995-1004); oracle and device define it: the earlier part of the query wins, whatever the blocks' spans
    ts_a = T0 + np.array([5, 6, 7], dtype=np.int64) * STEP            # part A's block starts later ...
    ts_b = T0 + np.array([1, 5, 6, 9], dtype=np.int64) * STEP         # ... than part B's block
    pa = _build(np.full(3, 4, np.uint64), ts_a, np.array([2, 2, 1], np.int64), f_int=np.array([10, 20, 30]))
    pb = _build(np.full(4, 4, np.uint64), ts_b, np.array([2, 2, 3, 2], np.int64), f_int=np.array([100, 200, 300, 400]))
    for parts, want in (([pa, pb], 100 + 10 + 300 + 30 + 400), ([pb, pa], 100 + 200 + 300 + 30 + 400)):
        r = O.run_query(O.Query(parts, [4], [("calls", O.AGG_SUM), ("calls", O.AGG_COUNT)]))
        assert r.val_i64[0].tolist() == [want, 5]


def test_block_selection_part_iter_test_go():
    # banyand/measure/part_iter_test.go:35-163 Test_partIter_nextBlock: which blocks a sorted series list selects
    from tests.helpers import PART_ITER_CASES, part_iter_fixture
    part = part_iter_fixture()
    for sids, want in PART_ITER_CASES:
        r = O.run_query(O.Query([part], sids, [("intField", O.AGG_COUNT)], groups=np.arange(len(sids), dtype=np.int32), n_groups=max(len(sids), 1),
                                tmin=1, tmax=220))
        assert r.blocks_scanned == len(want) and r.rows_scanned == 2 * len(want)
        assert [sids[g] for g in r.group_id.tolist()] == want and r.rows.tolist() == [2] * len(want)


def _model_keyed(sids, ts, f_int, f_flt, region, usid, groups, tmin, tmax, pred_code=None, code=None):
    """Insertion-ordered group-by on (group of the series, region value): aggregation.go:193-254."""
    order = np.lexsort((ts, sids))
    gmap = {int(s): int(g) for s, g in zip(usid, groups)}
    out = {}
    for i in order:
        if int(sids[i]) not in gmap or not (tmin <= ts[i] <= tmax):
            continue
        if pred_code is not None and code[i] != pred_code:
            continue
        rv = region[i] if region[i] is not None else b""
        k = (gmap[int(sids[i])], rv)
        e = out.setdefault(k, {"rows": 0, "si": 0, "sf": 0.0, "mx": -np.inf})
        e["rows"] += 1
        e["si"] += int(f_int[i])
        e["sf"] += float(f_flt[i])
        e["mx"] = max(e["mx"], float(f_flt[i]))
    return out


def test_group_by_row_tag_insertion_order():
    # a12: the key is a stored tag, so it changes from row to row inside a series
    sids, ts, ver, f_int, f_flt, region, code, _ = _synthetic(n_series=9, n_pts=400, seed=5)
    region = [None if (i % 53 == 0) else (b"" if i % 47 == 0 else r) for i, r in enumerate(region)]  # nil and "" are one key
    part = _build(sids, ts, ver, f_int, f_flt, region, code)
    usid = np.unique(sids)
    groups = (np.arange(usid.size) % 2).astype(np.int32)
    tmin, tmax = T0 + 10 * STEP, T0 + 350 * STEP
    aggs = [("calls", O.AGG_SUM), ("latency", O.AGG_SUM), ("latency", O.AGG_MAX), ("calls", O.AGG_COUNT)]
    q = O.Query([part], usid, aggs, groups=groups, n_groups=2, tmin=tmin, tmax=tmax,
                preds=[O.Pred("default", "code", O.OP_EQ, 200)], group_key=("default", "region"))
    r = O.run_query(q)
    want = _model_keyed(sids, ts, f_int, f_flt, region, usid, groups, tmin, tmax, pred_code=200, code=code)
    assert [(int(g), k) for g, k in zip(r.group_id, r.key)] == list(want.keys())
    for i, e in enumerate(want.values()):
        assert r.rows[i] == e["rows"] and r.val_i64[i, 0] == e["si"] and r.val_i64[i, 3] == e["rows"]
        assert abs(r.val_f64[i, 1] - e["sf"]) <= 1e-9 * abs(e["sf"]) and r.val_f64[i, 2] == e["mx"]
    assert (b"" in r.key) and len(set(zip(r.group_id.tolist(), r.key))) == len(r.key)
    # Top-N over the composite groups; ties go to the group inserted first (top.go:62-76)
    qt = O.Query([part], usid, [("calls", O.AGG_COUNT)], groups=groups, n_groups=2, tmin=tmin, tmax=tmax,
                 group_key=("default", "region"), top_n=5, top_agg=0, top_desc=True)
    rt = O.run_query(qt)
    full = _model_keyed(sids, ts, f_int, f_flt, region, usid, groups, tmin, tmax)
    ranked = sorted(enumerate(full.items()), key=lambda t: (-t[1][1]["rows"], t[0]))[:5]
    assert [(int(g), k) for g, k in zip(rt.group_id, rt.key)] == [kv[0] for _, kv in ranked]
    assert rt.val_i64[:, 0].tolist() == [kv[1]["rows"] for _, kv in ranked]


def test_group_by_row_tag_rejects_numeric_tag():
    sids, ts, ver, f_int, f_flt, region, code, _ = _synthetic(n_series=2, n_pts=20)
    part = _build(sids, ts, ver, f_int, f_flt, region, code)
    with pytest.raises(RuntimeError, match="group key"):
        O.run_query(O.Query([part], np.unique(sids), [("calls", O.AGG_SUM)], group_key=("default", "code")))
