#!/bin/bash
# Round 2, GPU call J (1 GPU): the group-key (a12) tests, then the whole GPU suite and a short bench (regression check of the lanes).
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
echo "== group-key tests"
timeout 600 python -m pytest tests -m gpu -q -x -k "group_by_stored_tag" 2>&1 | tail -30 | tee $OUT/${TAG}_keyed.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-700
tail -3 $OUT/${TAG}_bench.err
