#!/usr/bin/env python
"""bench_configs.py -- the larger BASELINE.json configurations, as extra measurements next to bench.py.

  C3  1e9 datapoints (10k series x 100k points, 4 float64 fields): GROUP BY service_id (1000 services x 10 series),
      sum(latency) + count(latency), Top 100 by the sum, one GPU
  C4  the same data sharded by series range over the ranks (torchrun, N = 2/4/8): STRONG scaling -- every rank scans
      its shard into a partial table, one NCCL all-gather, rank 0 combines in rank order and finalises

    python tools/bench_configs.py --steps 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --steps 10

Prints one JSON line (rank 0).  Algorithmic bytes per scanned datapoint for this query: 8 (the field)
+ 4 (group id) = 12 B; no row predicate, full time range, so the timestamp and tag pages are never read.
The query result is checked by its invariants: total count == datapoints, services' counts all equal, Top-100 sorted.
"""
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

B_ALG_C3 = 12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--series", type=int, default=10_000)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--services", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = B.load_pkg()
    ctx = pkg.Context(device=local_rank)
    per = args.series // world
    sid0 = 1 + rank * per
    n_mine = per if rank < world - 1 else args.series - per * (world - 1)
    t0 = time.perf_counter()
    img = B.make_part(pkg, n_mine, args.points, sid0, 0xB200 + rank)
    t_gen = time.perf_counter() - t0
    files = img.files()
    t0 = time.perf_counter()
    h = ctx.register_part(1 + rank, files)
    t_reg = time.perf_counter() - t0
    info = ctx.part_info(h)
    del img
    sids = np.arange(sid0, sid0 + n_mine, dtype=np.uint64)       # a rank resolves the series of its own shard
    groups = ((sids - 1) % args.services).astype(np.int32)          # service_id of a series: comes from the index, not the part
    q = pkg.Query(parts=[h], series_ids=sids, aggs=[("latency", pkg.AGG_SUM), ("latency", pkg.AGG_COUNT)], series_group=groups,
                  n_groups=args.services, top_n=100, top_agg=0, top_desc=True)
    pq = ctx.prepare(q)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stats = []
    if world == 1:
        def step(want_stats=True):
            r = ctx.scan_agg(pq)
            stats.append(r.stats)
            return r
    else:
        lay = ctx.partials_layout(q)
        words = lay["total_bytes"] // 8
        table = torch.zeros(words, dtype=torch.float64, device="cuda")
        gathered = torch.zeros(world * words, dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream

        def step(want_stats=False):
            st = ctx.scan_partials(pq, table.data_ptr(), lay["total_bytes"], stream, want_stats=want_stats)
            if st is not None:
                stats.append(st)
            dist.all_gather_into_tensor(gathered, table)
            if rank == 0:
                ctx.partials_combine(pq, gathered.data_ptr(), world, lay["total_bytes"], stream)
                return ctx.reduce_finalize(pq, gathered.data_ptr(), lay["total_bytes"], stream)
            torch.cuda.current_stream().synchronize()
            return None

    for _ in range(max(args.warmup, 3)):
        step()
    stats.clear()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = step()
    barrier()
    dt = time.perf_counter() - t
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        for _ in range(args.steps):
            step(want_stats=True)
        barrier()
    rows_step = stats[-1].rows_scanned
    scan_ms = float(np.mean([s.scan_kernel_ms for s in stats]))
    dev_ms = float(np.mean([s.device_ms for s in stats]))
    page_bytes = stats[-1].page_bytes
    total_rows = float(rows_step)
    if world > 1:
        tt = torch.tensor([dt, 0.0, scan_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sm = torch.tensor([float(rows_step)], dtype=torch.float64, device="cuda")
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, scan_ms, total_rows = float(tt[0]), float(tt[2]), float(sm[0])
    if rank != 0:
        dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = rows_step * B_ALG_C3 / (scan_ms * 1e-3) / 1e9
    sums, counts = last.val_f64[:, 0], last.val_i64[:, 1]
    checks = {"rows_returned": int(last.group_id.size), "top_sorted_desc": bool((np.diff(sums) <= 0).all()),
              "all_counts_equal": bool((counts == counts[0]).all()),
              "count_per_service_expected": int(counts[0]) == (args.series // args.services) * args.points}
    out = {"metric": "measure datapoints scanned+aggregated/sec", "value": total_rows * args.steps / dt, "unit": "datapoints/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong" if world > 1 else "n/a", "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.series * args.points:.0e} datapoints ({args.series} series x {args.points} points, 4 float64 fields), "
                                  f"GROUP BY service_id ({args.services} services), sum(latency)+count(latency), Top 100 desc by the sum"
                                  + (f"; series sharded over {world} ranks" if world > 1 else "")},
           "datapoints_per_step": total_rows, "device_ms_per_step_rank0": dev_ms, "scan_kernel_ms_max_rank": scan_ms,
           "roofline": {"bound": "hbm", "kernel": "scan_blocks_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_datapoint": B_ALG_C3, "encoded_page_bytes_per_launch_rank0": int(page_bytes),
                        "encoded_GBps_rank0": page_bytes / (scan_ms * 1e-3) / 1e9, "note": "per rank; rank 0's shard"},
           "clocks": clocks, "admission_rank0": {"generate_s": t_gen, "register_s": t_reg, **info},
           "result_checks": checks, "top3": [[int(g), float(s)] for g, s in zip(last.group_id[:3], sums[:3])]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
