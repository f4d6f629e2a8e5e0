#!/bin/bash
# Round 2, GPU call Y (1 GPU): launch list of the kernels behind the keyed scan and the page encoder; last whole-suite run of the round.
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_keyed_encode_launches.csv \
    python tools/probe_new_kernels.py > $OUT/${TAG}_probe.log 2>&1
tail -3 $OUT/${TAG}_probe.log | cut -c1-200
grep -E "key_|permute_table|encode_pages|gather_pages|index_" $OUT/${TAG}_keyed_encode_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -20 | tee $OUT/${TAG}_pytest.log
