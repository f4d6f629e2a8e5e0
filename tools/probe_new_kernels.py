#!/usr/bin/env python
"""probe_new_kernels.py -- one keyed query (bydb_scan_agg_keyed) and one page-encoder call (bydb_encode_pages) on a 1e8-datapoint
part, for an ncu launch list of the kernels behind them:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/probe_new_kernels.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    import torch
    torch.cuda.set_device(0)
    pkg = B.load_pkg()
    img = B.make_part(pkg, 1000, 100_000, 1)
    sids = np.arange(1, 1001, dtype=np.uint64)
    ctx = pkg.Context(device=0)
    h = ctx.register_part(1, img.files())
    q = pkg.Query(parts=[h], series_ids=sids, aggs=[("latency", pkg.AGG_SUM), ("latency", pkg.AGG_COUNT)])
    r = ctx.scan_agg_keyed(q, "default", "region")
    print("keyed:", [k.decode() for k in r.key], int(r.rows.sum()), "rows,", r.stats.kernel_launches, "launches")
    rng = np.random.default_rng(1)
    vals = np.round(25 + rng.normal(0, 5, 8192 * 1200), 2)
    pages, ms = ctx.encode_pages(vals, np.full(1200, 8192, dtype=np.uint32))
    print("encode:", sum(len(p) for p in pages if p is not None), "bytes,", ms, "ms")
    ctx.release_part(h)
    ctx.close()


if __name__ == "__main__":
    main()
