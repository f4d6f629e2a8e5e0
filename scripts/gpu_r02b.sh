#!/bin/bash
# Round 2, GPU call B: tests (new: SWAR path, C5 shape, HBM budget, scan_reduce, operator), 1e9 bench with the SWAR kernel, ncu.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -60 | tee $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== bench (b200 arm, 1e9)"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-2500
tail -5 $OUT/${TAG}_bench.err
echo "== ncu launch list (1e9, C3 query)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
grep -E "scan_blocks|select_rows|series_reduce|group_reduce|plan_blocks" $OUT/${TAG}_launches.csv | tail -8 | cut -d, -f5,15
echo "== ncu full capture of scan_blocks_kernel (C3, SWAR path)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_blocks -s 2 -c 1 -o $OUT/${TAG}_scan \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
