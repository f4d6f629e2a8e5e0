#!/usr/bin/env python
"""bench_encode.py -- write side: numeric field pages encoded on the device (bydb_encode_pages) against the C port of
the reference writer (oracle/part.c ob_column_encode) on the same blocks.

    python tools/bench_encode.py [--values 100000000] [--rows 8192]

Prints one JSON line: values/s through the C ABI (host values in, page bytes out), the kernels' own time, encoded bytes per value,
the CPU port's values/s on a bounded sample (one thread), and whether the sample's pages are byte-identical."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--values", type=int, default=100_000_000)
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-blocks", type=int, default=200)
    args = ap.parse_args()
    pkg = B.load_pkg()
    from oracle import oracle as O
    rng = np.random.default_rng(0xF4)
    n = args.values - args.values % args.rows
    out = {}
    for name, vals in (("latency float64 (2 decimals)", np.round(25 + rng.normal(0, 5, n), 2)),
                       ("fluctuating int64", 25 + np.cumsum(rng.integers(-5, 6, n)).astype(np.int64))):
        rows = np.full(n // args.rows, args.rows, dtype=np.uint32)
        ctx = pkg.Context(device=0)
        pages, ms = ctx.encode_pages(vals, rows)   # warm-up (module load, pool growth)
        wall, dev = [], []
        for _ in range(args.steps):
            t = time.perf_counter()
            pages, ms = ctx.encode_pages(vals, rows)
            wall.append(time.perf_counter() - t)
            dev.append(ms)
        ctx.close()
        nbytes = sum(len(p) for p in pages if p is not None)
        k = min(args.cpu_blocks, len(pages))
        t = time.perf_counter()
        same = True
        for b in range(k):
            blk = vals[b * args.rows:(b + 1) * args.rows]
            if blk.dtype == np.float64:
                raw, vt = blk.astype(">f8").tobytes(), O.VT_FLOAT64
            else:
                raw, vt = (blk.view(np.uint64) ^ np.uint64(1 << 63)).astype(">u8").tobytes(), O.VT_INT64
            same = same and O.column_encode(vt, [raw[8 * i:8 * i + 8] for i in range(blk.size)]) == pages[b]
        cpu_s = time.perf_counter() - t
        out[name] = {"values": int(n), "blocks": int(rows.size), "device_ms": float(np.median(dev)), "values_per_s_kernels": n / (np.median(dev) * 1e-3),
                     "wall_ms_host_in_host_out": float(np.median(wall) * 1e3), "values_per_s_e2e": n / float(np.median(wall)),
                     "encoded_bytes_per_value": nbytes / n, "blocks_left_to_cpu": int(sum(p is None for p in pages)),
                     "cpu_port_values_per_s_one_thread_incl_python_cells": k * args.rows / cpu_s, "cpu_sample_blocks": k, "sample_pages_identical": bool(same)}
    print(json.dumps({"tool": "bench_encode", "api": "bydb_encode_pages", "results": out}))


if __name__ == "__main__":
    main()
