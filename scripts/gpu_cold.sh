#!/bin/bash
# cold-path timeline under the default bench invocation (BYDB_TRACE: host-side phases of bydb_scan_agg_host on stderr)
mkdir -p gpurun_out
BYDB_TRACE=1 timeout 600 python bench.py 2>gpurun_out/cold_trace.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print({x:e[x] for x in ['value','ms_per_step','scan_kernel_ms','device_ms','d2h_bytes_per_step']}, d['ms_per_step'])"
grep "bydb cold" gpurun_out/cold_trace.err | tail -12
nproc; uptime
