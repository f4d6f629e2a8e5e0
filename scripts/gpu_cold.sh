#!/bin/bash
# cold-path slices experiment: e2e leg of bench.py for several BYDB_COLD_SLICES values
for k in "$@"; do
  BYDB_COLD_SLICES=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('K=$k', {x:e[x] for x in ['ms_per_step','scan_kernel_ms','device_ms']}, d['ms_per_step'])"
done
