#!/bin/bash
# one ncu --set full capture of the fast-lane scan kernel of the bench query (usage: bash scripts/gpu_ncu.sh tag)
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_blocks -s 2 -c 1 -o $OUT/${TAG}_scan -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
