#!/bin/bash
# Round 2, GPU call P (1 GPU): whole suite twice (the shared-device collective failed only inside the whole suite), bench short.
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2; do
echo "== pytest -m gpu (run $i)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  |Thread 0x|File \"/|Current thread" | head -60 | tee $OUT/${TAG}_pytest_$i.log
done
