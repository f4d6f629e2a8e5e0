#!/bin/bash
# cold-path check: gpu tests, then the e2e leg of bench.py with BYDB_TRACE (host-side phases of bydb_scan_agg_host on stderr)
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1; tail -2 gpurun_out/quick_pytest.log
BYDB_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/cold_trace.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print({x:e[x] for x in ['value','ms_per_step','scan_kernel_ms','device_ms']}, d['ms_per_step'])"
grep "bydb cold" gpurun_out/cold_trace.err | tail -14
