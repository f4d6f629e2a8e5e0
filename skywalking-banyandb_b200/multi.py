"""Multi-GPU reduce of the per-rank partial tables.

One process per GPU.  Every rank scans its own parts into a partial table (bydb_scan_partials); the
tables are combined with at most four tiny collectives over contiguous sub-ranges of ONE buffer
(layout: bydb_partials_layout in include/bydb_gpu.h):

    float64 SUM over [sum_f64]                     float64 MAX over [max_f64 | negmin_f64]
    int64   SUM over [sum_i64 | cnt | rows]        int64   MAX over [max_i64 | notmin_i64 | coltype]

which replaces the liaison's gather + reduceAccumulator.Combine
(pkg/query/logical/measure/measure_plan_aggregation.go:96-124).  min is carried as max of the negation
(float) / of the bitwise complement (int64) so that one MAX covers both.  Works on any backend
(NCCL on GPUs; gloo in the CPU tests).

bench.py uses the other reduce the C ABI offers: ONE all-gather of the tables followed by
bydb_partials_combine (rank-ordered, hence deterministic float sums) and bydb_reduce_finalize.
"""
from __future__ import annotations

from typing import Dict


def allreduce_partial_table(table_f64, layout: Dict[str, int], dist, need_minmax: bool = True, need_int: bool = True) -> int:
    """All-reduces the partial table in place. ``table_f64`` is a 1-D float64 tensor of total_bytes/8
    elements; the int64 ranges are reduced through an int64 view of the same storage.  Returns the
    number of collectives issued."""
    import torch

    ti64 = table_f64.view(torch.int64)
    n = 0
    a, k = layout["off_sum_f64"] // 8, layout["n_sum_f64"]
    dist.all_reduce(table_f64[a:a + k], op=dist.ReduceOp.SUM)
    n += 1
    if need_minmax:
        a, k = layout["off_max_f64"] // 8, layout["n_max_f64"]
        dist.all_reduce(table_f64[a:a + k], op=dist.ReduceOp.MAX)
        n += 1
    # counts and rows live in the int64 SUM range, the column types in the int64 MAX range: always needed
    a, k = layout["off_sum_i64"] // 8, layout["n_sum_i64"]
    dist.all_reduce(ti64[a:a + k], op=dist.ReduceOp.SUM)
    n += 1
    a, k = layout["off_max_i64"] // 8, layout["n_max_i64"]
    if not (need_minmax or need_int):
        k_all = k
        f = k_all - 2 * layout["n_sum_f64"]          # only the trailing coltype[F] words
        a, k = a + 2 * layout["n_sum_f64"], f
    dist.all_reduce(ti64[a:a + k], op=dist.ReduceOp.MAX)
    n += 1
    return n
