#!/bin/bash
# usage: bash scripts/gpu_variants.sh lib1.so lib2.so ...   (paths relative to skywalking-banyandb_b200/): bench kernel time per variant
for v in "$@"; do
  BYDB_GPU_LIB=$PWD/skywalking-banyandb_b200/$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', {k:d[k] for k in ['ms_per_step','device_ms_per_step','scan_kernel_ms']}, d['result'])"
done
