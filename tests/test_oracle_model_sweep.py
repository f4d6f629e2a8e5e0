"""Randomised check of the oracle's query path against a brute-force model written independently of its merge code:
all rows of 1-3 overlapping parts in a dict keyed by (series, timestamp) -> highest version (earlier part on equal versions),
row predicates with nil cells, time range, groups, the five functions with the reference's MEAN rules and the row-order
float sum.  Values are read back from each part first, because the decimal float pages are lossy for 17-digit inputs."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import STEP, T0, build_part

OPS = {O.OP_EQ: lambda h, c: h and c == 0, O.OP_NE: lambda h, c: (not h) or c != 0, O.OP_LT: lambda h, c: h and c < 0,
       O.OP_LE: lambda h, c: h and c <= 0, O.OP_GT: lambda h, c: h and c > 0, O.OP_GE: lambda h, c: h and c >= 0}
AGGS = [("i", O.AGG_SUM), ("i", O.AGG_COUNT), ("i", O.AGG_MIN), ("i", O.AGG_MAX), ("i", O.AGG_MEAN),
        ("f", O.AGG_SUM), ("f", O.AGG_COUNT), ("f", O.AGG_MIN), ("f", O.AGG_MAX), ("f", O.AGG_MEAN)]


def random_case(seed):
    """-> (parts, best-row dict, query kwargs): shared with the device sweep in test_gpu_parity.py"""
    rng = np.random.default_rng(9000 + seed)
    nparts, nser = int(rng.integers(1, 4)), int(rng.integers(1, 7))
    best, parts = {}, []
    for pi in range(nparts):
        rows = []
        for s in range(1, nser + 1):
            if rng.random() < 0.2:
                continue
            npts = int(rng.choice([1, 3, 40, 300, 9000] if seed % 4 == 0 else [1, 3, 40, 300]))
            for t in np.sort(rng.choice(np.arange(0, 12000), size=npts, replace=False)):
                rows.append((s, int(T0 + t * STEP), int(rng.integers(0, 4))))
        rows = sorted(rows) or [(1, int(T0), 1)]
        n = len(rows)
        sid = np.array([r[0] for r in rows], dtype=np.uint64)
        ts = np.array([r[1] for r in rows], dtype=np.int64)
        ver = np.array([r[2] for r in rows], dtype=np.int64)
        fi, fin = rng.integers(-1000, 1000, n), (rng.random(n) < 0.15).astype(np.uint8)
        ff = np.round(rng.normal(10, 5, n), 2) if rng.random() < 0.7 else rng.standard_normal(n) * 10.0 ** rng.integers(-4, 5, n)
        ffn = (rng.random(n) < 0.1).astype(np.uint8)
        tagv = [None if rng.random() < 0.1 else (b"r%d" % rng.integers(0, 4)) for _ in range(n)]
        code, coden = rng.integers(0, 5, n), (rng.random(n) < 0.1).astype(np.uint8)
        part = build_part(sid, ts, ver, [("i", O.VT_INT64, fi, fin if fin.any() else None), ("f", O.VT_FLOAT64, ff, ffn if ffn.any() else None)],
                          [("default", [("region", O.VT_STR, tagv, None), ("code", O.VT_INT64, code, coden if coden.any() else None)])])
        parts.append(part)
        stored = O.scan_rows(O.Query([part], sorted(set(int(x) for x in sid)), [("i", O.AGG_SUM), ("f", O.AGG_SUM)]))["fields"][1][2]
        for k in range(n):
            key, cand = (int(sid[k]), int(ts[k])), (int(ver[k]), -pi)
            row = dict(i=None if fin[k] else int(fi[k]), f=None if ffn[k] else float(stored[k]), region=tagv[k], code=None if coden[k] else int(code[k]))
            if key not in best or cand > best[key][0]:
                best[key] = (cand, row)
    tmin, tmax = int(T0 + int(rng.integers(0, 6000)) * STEP), int(T0 + int(rng.integers(6000, 12000)) * STEP)
    preds = []
    if rng.random() < 0.6:
        preds.append(("region", int(rng.choice(list(OPS))), b"r%d" % rng.integers(0, 4)))
    if rng.random() < 0.4:
        preds.append(("code", int(rng.choice(list(OPS))), int(rng.integers(0, 5))))
    G = int(rng.integers(1, 4))
    groups = [int(rng.integers(0, G)) for _ in range(nser)]
    return parts, best, dict(nser=nser, groups=groups, G=G, tmin=tmin, tmax=tmax, preds=preds, threads=int(rng.integers(1, 4)))


def case_query(parts, kw):
    return O.Query(parts, list(range(1, kw["nser"] + 1)), AGGS, groups=np.array(kw["groups"], dtype=np.int32), n_groups=kw["G"], tmin=kw["tmin"],
                   tmax=kw["tmax"], preds=[O.Pred("default", t, op, v) for t, op, v in kw["preds"]], threads=kw["threads"])


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_brute_force_model(seed):
    parts, best, kw = random_case(seed)
    groups, tmin, tmax, preds = kw["groups"], kw["tmin"], kw["tmax"], kw["preds"]
    res = O.run_query(case_query(parts, kw))
    exp = {}
    for (s, t), (_, row) in sorted(best.items()):
        if t < tmin or t > tmax:
            continue
        if not all(OPS[op](row[tag] is not None, 0 if row[tag] is None else (row[tag] > lit) - (row[tag] < lit)) for tag, op, lit in preds):
            continue
        e = exp.setdefault(groups[s - 1], dict(rows=0, i=[], f=[]))
        e["rows"] += 1
        if row["i"] is not None:
            e["i"].append(row["i"])
        if row["f"] is not None:
            e["f"].append(row["f"])
    got = {int(g): k for k, g in enumerate(res.group_id.tolist())}
    assert sorted(got) == sorted(exp)
    for g, e in exp.items():
        k, iv, fv = got[g], e["i"], e["f"]
        assert int(res.rows[k]) == e["rows"] and int(res.val_i64[k, 1]) == len(iv) and int(res.val_i64[k, 6]) == len(fv)
        if iv:
            mean = sum(iv) // len(iv) if sum(iv) >= 0 else -((-sum(iv)) // len(iv))          # Go integer division truncates
            assert [int(res.val_i64[k, c]) for c in (0, 2, 3, 4)] == [sum(iv), min(iv), max(iv), max(mean, 1)]
        if fv:
            acc = 0.0
            for x in fv:                                                                     # function.go:133-135: row order
                acc += x
            mean = acc / len(fv)
            assert [float(res.val_f64[k, c]) for c in (5, 7, 8, 9)] == [acc, min(fv), max(fv), 1.0 if mean < 1 else mean]
