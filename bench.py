#!/usr/bin/env python
"""bench.py -- the measure-query hot path on B200: scanned datapoints/s and achieved HBM GB/s.

A "step" is one pass of the hot path (block selection -> decode -> time/tag filter -> aggregate)
over the synthetic measure of BASELINE.json configs[1]:
    1e8 datapoints (1k series x 100k points), 4 float64 fields, time range (middle 50%) AND
    region == "r3", avg(latency) + max(walk), scalar result.
``value``   datapoints scanned+aggregated per second with the parts already resident in HBM.
``e2e``     the same metric through the host-buffer entry point (bydb_scan_agg_host): part files in
            pinned host memory are uploaded, scanned and the result read back inside the timed region.
``roofline``  algorithmic bytes (25 B per scanned datapoint for this query) divided by the
            scan kernel's CUDA-event time, against the measured HBM copy bandwidth.
``cpu_baseline`` the oracle (C port of the reference's Go path; the Go reference cannot be built here)
            on the host cores over a bounded sample of the same part.
``--impl reference`` times that CPU port alone (all host threads) and prints the same JSON shape.

N > 1 (torchrun): every rank owns one part of the same shape (weak scaling, series-disjoint), runs the
scan into a partial table on its GPU, the tiny tables are exchanged with ONE NCCL all-gather and rank 0
combines them in rank order (deterministic) and finalises.  No other data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T0 = 1_700_000_000_000_000_000
STEP = 60_000_000_000
B_ALG = 25  # bytes per scanned datapoint for this query: 8 (timestamp) + 1 (dictionary tag) + 2 x 8 (fields)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--series", type=int, default=1000)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graph", action="store_true", help="also time the prepared-query path (one captured CUDA graph per step)")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def load_pkg():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libbydbgpu.so")):
        ge.build()
    return ge.load_package()


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of one scan_blocks launch of this query, from the committed
    ncu --set full capture (profiles/traffic.json names it); None when there is none."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        return None


def make_part(pkg, n_series, n_points, sid0, seed):
    from importlib import import_module
    S = import_module("bydb_b200.synth")
    fields = [("latency", S.F_LATENCY), ("walk", S.F_WALK3), ("ints", S.F_INT1000), ("uniform", S.F_UNIFORM)]
    return S.synth_part(n_series, n_points, fields, sid0=sid0, sid_step=1, t0=T0, t_step=STEP, region_values=8, region_run=16,
                        seed=seed)


def query_of(pkg, handles, sids, n_points):
    tmin = T0 + (n_points // 4) * STEP
    tmax = T0 + (3 * n_points // 4) * STEP
    return pkg.Query(parts=handles, series_ids=sids, aggs=[("latency", pkg.AGG_MEAN), ("walk", pkg.AGG_MAX)], tmin=tmin, tmax=tmax,
                     preds=[pkg.Pred("default", "region", pkg.OP_EQ, b"r3")])


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(files, sids_all, n_points, threads, target_seconds, per_thread_partials=False):
    """Datapoints/s of the oracle on a bounded sample.  Default = reference-shaped (decode pool + single-threaded
    merge/fold, like mergeBatch/Consume); per_thread_partials = the optimistic all-core variant of SURVEY.md 8(d)."""
    from oracle import oracle as O
    part = O.Part.open({k: bytes(v) for k, v in files.items()})
    tmin = T0 + (n_points // 4) * STEP
    tmax = T0 + (3 * n_points // 4) * STEP

    def run(nser):
        q = O.Query([part], sids_all[:nser], [("latency", O.AGG_MEAN), ("walk", O.AGG_MAX)], tmin=tmin, tmax=tmax,
                    preds=[O.Pred("default", "region", O.OP_EQ, b"r3")], threads=threads, per_thread_partials=per_thread_partials)
        t = time.perf_counter()
        r = O.run_query(q)
        return time.perf_counter() - t, r

    probe = max(1, min(len(sids_all), 8))
    dt, r = run(probe)
    rate = r.rows_scanned / max(dt, 1e-9)
    nser = int(max(probe, min(len(sids_all), target_seconds * rate / max(r.rows_scanned / probe, 1))))
    dt, r = run(nser)
    return r.rows_scanned / dt, nser, r, dt


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_series, n_points = args.series, args.points
    workload = f"{n_series * n_points:.0e} datapoints ({n_series} series x {n_points} points), 4 float64 fields, " \
               f"time range (middle 50%) AND region==\"r3\", avg(latency)+max(walk)"
    cfg = {"workload": workload, "n_series": n_series, "n_points": n_points, "query": "mean(latency), max(walk)",
           "timing": "inputs larger than L2 (no flush needed): ~200 MB of encoded pages per step vs 126 MB L2"}
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm: the CPU port only
    if args.impl == "reference":
        if rank != 0:
            return
        pkg = load_pkg()
        img = make_part(pkg, n_series, n_points, 1, 0xB200)
        sids = np.arange(1, n_series + 1, dtype=np.uint64)
        files = img.files()
        per_step_target = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
        rate0, nser, _, _ = cpu_port_rate(files, sids, n_points, cores, per_step_target)
        from oracle import oracle as O
        part = O.Part.open({k: bytes(v) for k, v in files.items()})
        q = O.Query([part], sids[:nser], [("latency", O.AGG_MEAN), ("walk", O.AGG_MAX)], tmin=T0 + (n_points // 4) * STEP,
                    tmax=T0 + (3 * n_points // 4) * STEP, preds=[O.Pred("default", "region", O.OP_EQ, b"r3")], threads=cores)
        for _ in range(args.warmup):
            O.run_query(q)
        t = time.perf_counter()
        rows = 0
        for _ in range(args.steps):
            rows += O.run_query(q).rows_scanned
        dt = time.perf_counter() - t
        val = rows / dt
        sample = f"{nser} of {n_series} series of the same part per step ({rows // max(args.steps, 1)} datapoints/step)"
        print(json.dumps({"metric": "measure datapoints scanned+aggregated/sec", "value": val, "unit": "datapoints/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": cfg,
                          "cpu_baseline": {"value": val, "unit": "datapoints/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the measure scan path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = load_pkg()
    ctx = pkg.Context(device=local_rank)
    sid0 = 1 + rank * n_series
    img = make_part(pkg, n_series, n_points, sid0, 0xB200 + rank)
    files = img.files()
    t_reg = time.perf_counter()
    h = ctx.register_part(1 + rank, files)
    admission = {"register_ms": (time.perf_counter() - t_reg) * 1e3, **ctx.part_info(h),
                 "note": "one-time per part: upload to HBM, host parse of the block index, device unpack of the fallback pages "
                         "(the `uniform` field is full-precision float64 = zstd-compressed EncodeTypePlain pages)"}
    # every rank resolves the series of ITS shard (a data node's index lookup returns local series only); the scalar /
    # group layout of the partial table is the same on all ranks, so the tables combine
    sids = np.arange(sid0, sid0 + n_series, dtype=np.uint64)
    q = query_of(pkg, [h], sids, n_points)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stats_acc = []
    pq = ctx.prepare(q)          # the query is marshalled to the C struct once, like a cgo caller would hold it
    phase = {"scan_enqueue": 0.0, "all_gather_enqueue": 0.0, "combine_finalize_sync": 0.0}
    if world == 1:
        def step():
            r = ctx.scan_agg(pq)
            stats_acc.append(r.stats)
            return r
    else:
        lay = ctx.partials_layout(q)
        words = lay["total_bytes"] // 8
        table = torch.zeros(words, dtype=torch.float64, device="cuda")
        gathered = torch.zeros(world * words, dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream

        def step(want_stats=False):
            # map: every rank scans its own parts into a partial table on its GPU.  Asynchronous: the scan, the
            # collective and the finalisation are enqueued back to back; the only host wait of the step is the
            # result read-back on rank 0 (a failing block travels in the table and fails reduce_finalize).
            t0 = time.perf_counter()
            st = ctx.scan_partials(pq, table.data_ptr(), lay["total_bytes"], stream, want_stats=want_stats)
            if st is not None:
                stats_acc.append(st)
            t1 = time.perf_counter()
            # reduce: ONE NCCL collective (all-gather of the tiny tables over NVLink), then a deterministic
            # rank-ordered combine + finalisation on rank 0's GPU
            dist.all_gather_into_tensor(gathered, table)
            t2 = time.perf_counter()
            res = None
            if rank == 0:
                ctx.partials_combine(pq, gathered.data_ptr(), world, lay["total_bytes"], stream)
                res = ctx.reduce_finalize(pq, gathered.data_ptr(), lay["total_bytes"], stream)
            else:
                torch.cuda.current_stream().synchronize()
            t3 = time.perf_counter()
            phase["scan_enqueue"] += t1 - t0
            phase["all_gather_enqueue"] += t2 - t1
            phase["combine_finalize_sync"] += t3 - t2
            return res

    for _ in range(max(args.warmup, 3)):
        step()
    stats_acc.clear()
    for k in phase:
        phase[k] = 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = step()
    barrier()
    dt = time.perf_counter() - t
    clocks = None  # the sampler keeps running through the e2e legs: the resident timed region alone is only a few ms long
    kernel_timing = "cuda events inside the timed steps"
    phase_timed = dict(phase)
    if world > 1:
        # the timed steps are asynchronous and carry no statistics: per-kernel device times (CUDA events inside
        # bydb_scan_partials) come from the same steps run once more with statistics on, outside the timed region
        for _ in range(args.steps):
            step(want_stats=True)
        barrier()
        kernel_timing = "cuda events in a second pass of the same steps (the timed steps are asynchronous, no statistics read-back)"
    rows_step = stats_acc[-1].rows_scanned
    scan_ms = float(np.mean([s.scan_kernel_ms for s in stats_acc]))
    dev_ms = float(np.mean([s.device_ms for s in stats_acc]))
    launches = int(sum(s.kernel_launches for s in stats_acc)) + (3 * args.steps if world > 1 and rank == 0 else 0)
    page_bytes = stats_acc[-1].page_bytes
    if world > 1:
        tt = torch.tensor([dt, float(rows_step)], dtype=torch.float64, device="cuda")
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt = float(mx[0])
        total_rows_step = float(sm[1])
    else:
        total_rows_step = float(rows_step)
    value = total_rows_step * args.steps / dt

    # ------------------------------------------------------------------ end to end: host buffers in, result out
    e2e = None
    if not args.no_e2e:
        pinned, keep_pinned = {}, []
        for k, v in files.items():
            tns = torch.empty(v.size + 256, dtype=torch.uint8, pin_memory=True)   # 256 B of readable slack after each image
            tns[:v.size].copy_(torch.from_numpy(np.ascontiguousarray(v)))
            keep_pinned.append(tns)
            pinned[k] = tns[:v.size].numpy()

        def e2e_leg(flags, steps):
            qh = query_of(pkg, [], sids, n_points)
            qh.flags = flags
            ctx.scan_agg_host([pinned], qh)
            barrier()
            t0 = time.perf_counter()
            st = None
            for _ in range(steps):
                st = ctx.scan_agg_host([pinned], qh).stats
            barrier()
            d = time.perf_counter() - t0
            if world > 1:
                m = torch.tensor([d], dtype=torch.float64, device="cuda")
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                d = float(m[0])
            return {"value": total_rows_step * steps / d, "unit": "datapoints/s", "h2d_bytes_per_step": int(st.h2d_bytes),
                    "d2h_bytes_per_step": int(st.d2h_bytes), "ms_per_step": d / steps * 1e3, "steps": steps,
                    "scan_kernel_ms": st.scan_kernel_ms, "device_ms": st.device_ms}

        e2e_steps = max(3, min(args.steps, 10))
        from bydb_b200.capi import Q_HOST_ZERO_COPY
        staged = e2e_leg(0, max(3, e2e_steps // 2))
        staged["note"] = "bydb_scan_agg_host: every file of the part is copied from pinned host memory to HBM, scanned, result copied back"
        try:
            e2e = e2e_leg(Q_HOST_ZERO_COPY, e2e_steps)
            e2e["note"] = ("bydb_scan_agg_host(BYDB_Q_HOST_ZERO_COPY): part files stay in pinned host memory; every step parses the "
                           "block index, uploads the block directory and the kernels pull exactly the pages the query touches over "
                           "PCIe (h2d = directory + page bytes), result copied back"
                           + ("; per rank, no cross-rank reduce in this leg" if world > 1 else ""))
            e2e["staged_upload"] = staged
        except Exception as ex:  # keep the bench line alive: report the staged leg as e2e
            e2e = staged
            e2e["zero_copy_error"] = str(ex)[:200]

    if rank == 0:
        clocks = sampler.stop()
        clocks["window"] = "resident timed steps + e2e timed steps (100 ms sampling)"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = rows_step * B_ALG / (scan_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_blocks_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": int(rows_step * B_ALG), "kernel_ms": scan_ms,
                "encoded_page_bytes_per_launch": int(page_bytes), "encoded_GBps": page_bytes / (scan_ms * 1e-3) / 1e9,
                "traffic": traffic_from_profile(), "kernel_timing": kernel_timing}
    out = {"metric": "measure datapoints scanned+aggregated/sec", "value": value, "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": cfg, "datapoints_per_step": total_rows_step, "device_ms_per_step": dev_ms,
           "scan_kernel_ms": scan_ms, "blocks_slow_lane": int(stats_acc[-1].blocks_slow_lane), "slow_lane_reasons": int(stats_acc[-1].slow_lane_reasons), "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "e2e": e2e}
    out["part_admission"] = admission
    if world == 1:
        # the fallback (zstd) field is reported separately -- same predicate and range over `uniform`
        qf = pkg.Query(parts=[h], series_ids=sids, aggs=[("uniform", pkg.AGG_MEAN), ("uniform", pkg.AGG_MAX)], tmin=q.tmin, tmax=q.tmax,
                       preds=[pkg.Pred("default", "region", pkg.OP_EQ, b"r3")])
        pqf = ctx.prepare(qf)
        for _ in range(3):
            rf = ctx.scan_agg(pqf)
        barrier()
        t0 = time.perf_counter()
        nf = max(3, min(args.steps, 10))
        for _ in range(nf):
            rf = ctx.scan_agg(pqf)
        barrier()
        df = (time.perf_counter() - t0) / nf
        out["fallback_field_query"] = {"query": "mean(uniform), max(uniform), same range and predicate", "ms_per_step": df * 1e3,
                                       "value": rf.stats.rows_scanned / df, "unit": "datapoints/s", "scan_kernel_ms": rf.stats.scan_kernel_ms,
                                       "blocks_slow_lane": int(rf.stats.blocks_slow_lane), "mean": float(rf.val_f64[0, 0]), "max": float(rf.val_f64[0, 1]),
                                       "note": "raw-cell pages written at admission (unpack_kernels.cu), scanned by the general lane"}
    if world == 1 and args.graph:
        # the same query through bydb_query_prepare / bydb_scan_agg_prepared: step 1 ordinary, step 2 capture, then replays
        gq = ctx.prepare_graph(q)
        try:
            for _ in range(max(args.warmup, 3) + 2):
                rg = gq.run()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                rg = gq.run()
            barrier()
            dg = (time.perf_counter() - t0) / args.steps
            out["prepared_graph"] = {"ms_per_step": dg * 1e3, "value": rg.stats.rows_scanned / dg, "unit": "datapoints/s", "device_ms": rg.stats.device_ms,
                                     "same_result": bool(last is not None and rg.val_f64.tolist() == last.val_f64.tolist() and rg.rows.tolist() == last.rows.tolist()),
                                     "note": "bydb_scan_agg_prepared: the whole step replayed as one CUDA graph (one launch + one synchronisation)"}
        except Exception as ex:
            out["prepared_graph"] = {"error": str(ex)[:200]}
        finally:
            gq.close()
    if world > 1:
        out["host_phase_ms_per_step_rank0"] = {k: v / args.steps * 1e3 for k, v in phase_timed.items()}
    if last is not None:
        out["result"] = {"mean_latency": float(last.val_f64[0, 0]), "max_walk": float(last.val_f64[0, 1]), "rows_matched": int(last.rows[0])}
    if world == 1 and not args.no_cpu:
        rate, nser, r, cdt = cpu_port_rate(files, sids, n_points, cores, args.cpu_seconds)
        out["cpu_baseline"] = {"value": rate, "unit": "datapoints/s", "cores": cores, "kind": "port",
                               "sample": f"{nser} of {n_series} series of the same part ({r.rows_scanned} datapoints, {cdt:.1f} s); "
                                         "C port of the reference Go path: decode on a thread pool, single-threaded merge+fold"}
        try:  # the optimistic variant (every core folds its own partials), so the GPU is not compared against a strawman
            rate2, nser2, _, cdt2 = cpu_port_rate(files, sids, n_points, cores, min(args.cpu_seconds, 6.0), per_thread_partials=True)
            out["cpu_baseline"]["all_core_partials_variant"] = {"value": rate2, "unit": "datapoints/s", "cores": cores,
                                                                "sample": f"{nser2} of {n_series} series, {cdt2:.1f} s; per-thread partial aggregates"}
        except Exception as ex:
            out["cpu_baseline"]["all_core_partials_variant"] = {"error": str(ex)[:120]}
        if last is not None and nser == n_series:
            out["cpu_baseline"]["agrees_with_gpu"] = bool(abs(r.val_f64[0, 0] - last.val_f64[0, 0]) <= 1e-9 * abs(r.val_f64[0, 0])
                                                          and r.val_f64[0, 1] == last.val_f64[0, 1])
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
