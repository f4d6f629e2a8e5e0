#!/bin/bash
# Round 2, GPU call U (1 GPU): host-polled waits for ranks that share a device -- the whole suite three times.
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
echo "== pytest -m gpu (run $i)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 | grep -E "FAILED|ERROR|passed|failed|^E  |Thread 0x|File \"/|Current thread|line [0-9]+ in" | head -60 | tee $OUT/${TAG}_pytest_$i.log
done
