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
