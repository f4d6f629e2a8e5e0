"""N>1 host logic on CPU: two gloo ranks build the partial tables of their own (series-disjoint) parts,
combine them with allreduce_partial_table (the code bench.py / a multi-GPU node runs over NCCL) and must
obtain the table of the global query.  The per-rank tables come from the oracle here (no GPU in this
container); the layout comes from the product's bydb_partials_layout, which needs no device."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from tests.helpers import STEP, T0, grid

F64_NEG_INF = float("-inf")
I64_MIN = -(1 << 63)


def _rank_data(rank):
    rng = np.random.default_rng(100 + rank)
    sids, ts, ver = grid(6, 700, sid0=1 + rank * 6)
    lat = np.round(rng.normal(30, 8, sids.size), 2)
    calls = rng.integers(-100, 1000, sids.size)
    return sids, ts, ver, lat, calls


def _fill_table(layout, q: O.Query, G, fields, types):
    """numpy partial table (as float64 words) of one rank from oracle SUM/COUNT/MIN/MAX results."""
    tab = np.zeros(layout["total_bytes"] // 8, dtype=np.float64)
    ti = tab.view(np.int64)
    F = len(fields)
    GF = G * F
    o = {k: layout[k] // 8 for k in ("off_sum_f64", "off_max_f64", "off_sum_i64", "off_max_i64")}
    tab[o["off_max_f64"]:o["off_max_f64"] + 2 * GF] = F64_NEG_INF
    ti[o["off_max_i64"]:o["off_max_i64"] + 2 * GF] = I64_MIN
    aggs = []
    for f in fields:
        aggs += [(f, O.AGG_SUM), (f, O.AGG_COUNT), (f, O.AGG_MIN), (f, O.AGG_MAX)]
    q.aggs = aggs
    r = O.run_query(q)
    for row, g in enumerate(r.group_id.tolist()):
        ti[o["off_sum_i64"] + 2 * GF + g] = r.rows[row]
        for c, f in enumerate(fields):
            at = g * F + c
            cnt = int(r.val_i64[row, 4 * c + 1])
            ti[o["off_sum_i64"] + GF + at] = cnt
            if cnt == 0:
                continue
            if types[c] == O.VT_FLOAT64:
                tab[o["off_sum_f64"] + at] = r.val_f64[row, 4 * c]
                tab[o["off_max_f64"] + at] = r.val_f64[row, 4 * c + 3]
                tab[o["off_max_f64"] + GF + at] = -r.val_f64[row, 4 * c + 2]
            else:
                ti[o["off_sum_i64"] + at] = r.val_i64[row, 4 * c]
                ti[o["off_max_i64"] + at] = r.val_i64[row, 4 * c + 3]
                ti[o["off_max_i64"] + GF + at] = ~int(r.val_i64[row, 4 * c + 2])
    if r.group_id.size:
        for c in range(F):
            ti[o["off_max_i64"] + 2 * GF + c] = types[c]
    return tab


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    bydb = ge.load_package()
    from importlib import import_module
    multi = import_module("bydb_b200.multi")
    sids, ts, ver, lat, calls = _rank_data(rank)
    b = O.PartBuilder()
    b.append(sids, ts, ver, [("latency", O.VT_FLOAT64, lat, None), ("calls", O.VT_INT64, calls, None)])
    part = b.finish()
    all_sids = np.arange(1, 1 + 6 * world, dtype=np.uint64)
    groups = (np.arange(all_sids.size) % 4).astype(np.int32)
    gq = bydb.Query(parts=[], series_ids=all_sids, aggs=[("latency", bydb.AGG_MEAN), ("calls", bydb.AGG_MIN)], series_group=groups, n_groups=4,
                    tmin=T0 + 50 * STEP, tmax=T0 + 600 * STEP)
    # bydb_partials_layout is pure host logic: usable without a device or a context
    from bydb_b200.capi import _Layout, _mk_query, load_library
    import ctypes as C
    keep = []
    lay = _Layout()
    assert load_library().bydb_partials_layout(C.byref(_mk_query(gq, keep)), C.byref(lay)) == 0
    layout = {k: getattr(lay, k) for k, _ in _Layout._fields_}
    oq = O.Query([part], all_sids, [], groups=groups, n_groups=4, tmin=gq.tmin, tmax=gq.tmax)
    tab = torch.from_numpy(_fill_table(layout, oq, 4, ["latency", "calls"], [O.VT_FLOAT64, O.VT_INT64]))
    mine = tab.clone()
    n = multi.allreduce_partial_table(tab, layout, dist)
    assert n == 4
    # the other reduce (the one bench.py runs over NCCL): ONE all-gather of the tables, then the rank-ordered combine that
    # bydb_partials_combine performs on the device -- restated here in numpy: SUM ranges add in rank order, MAX ranges take the max
    gathered = torch.empty(world * mine.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, mine)
    tabs = gathered.numpy().reshape(world, -1)
    comb = tabs[0].copy()
    ci = comb.view(np.int64)
    for r in range(1, world):
        ti = tabs[r].view(np.int64)
        a, k = layout["off_sum_f64"] // 8, layout["n_sum_f64"]
        comb[a:a + k] += tabs[r][a:a + k]
        a, k = layout["off_max_f64"] // 8, layout["n_max_f64"]
        comb[a:a + k] = np.maximum(comb[a:a + k], tabs[r][a:a + k])
        a, k = layout["off_sum_i64"] // 8, layout["n_sum_i64"]
        ci[a:a + k] += ti[a:a + k]
        a, k = layout["off_max_i64"] // 8, layout["n_max_i64"]
        ci[a:a + k] = np.maximum(ci[a:a + k], ti[a:a + k])
    assert comb.view(np.int64).tolist() == tab.numpy().view(np.int64).tolist(), "all-gather + combine must equal the all-reduce result"
    if rank == 0:
        np.save(out, tab.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_global_query(tmp_path, bydb):
    world, port = 2, 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "table.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    # the same table from ONE oracle query over both ranks' parts
    parts = []
    for rank in range(world):
        sids, ts, ver, lat, calls = _rank_data(rank)
        b = O.PartBuilder()
        b.append(sids, ts, ver, [("latency", O.VT_FLOAT64, lat, None), ("calls", O.VT_INT64, calls, None)])
        parts.append(b.finish())
    all_sids = np.arange(1, 1 + 6 * world, dtype=np.uint64)
    groups = (np.arange(all_sids.size) % 4).astype(np.int32)
    from bydb_b200.capi import _Layout, _mk_query, load_library
    import ctypes as C
    gq = bydb.Query(parts=[], series_ids=all_sids, aggs=[("latency", bydb.AGG_MEAN), ("calls", bydb.AGG_MIN)], series_group=groups, n_groups=4)
    lay = _Layout()
    keep = []
    assert load_library().bydb_partials_layout(C.byref(_mk_query(gq, keep)), C.byref(lay)) == 0
    layout = {k: getattr(lay, k) for k, _ in _Layout._fields_}
    assert layout["n_sum_f64"] == 8 and layout["n_max_f64"] == 16 and layout["n_sum_i64"] == 20 and layout["n_max_i64"] == 18
    oq = O.Query(parts, all_sids, [], groups=groups, n_groups=4, tmin=T0 + 50 * STEP, tmax=T0 + 600 * STEP)
    want = _fill_table(layout, oq, 4, ["latency", "calls"], [O.VT_FLOAT64, O.VT_INT64])
    gi, wi = got.view(np.int64), want.view(np.int64)
    a, k = layout["off_sum_f64"] // 8, layout["n_sum_f64"]
    np.testing.assert_allclose(got[a:a + k], want[a:a + k], rtol=1e-12)          # float sums: order differs
    a, k = layout["off_max_f64"] // 8, layout["n_max_f64"]
    assert got[a:a + k].tolist() == want[a:a + k].tolist()                        # float min/max: exact
    a, k = layout["off_sum_i64"] // 8, layout["n_sum_i64"]
    assert gi[a:a + k].tolist() == wi[a:a + k].tolist()                           # int sums, counts, rows: exact
    a, k = layout["off_max_i64"] // 8, layout["n_max_i64"]
    assert gi[a:a + k].tolist() == wi[a:a + k].tolist()                           # int min/max, column types
