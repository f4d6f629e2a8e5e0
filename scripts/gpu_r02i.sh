#!/bin/bash
# Round 2, GPU call I (1 GPU): the round's reference run -- full GPU suite, smoke, both bench arms, C5 tool, ncu launch list
# (plain calls) and a full capture of the express lane.
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $OUT/${TAG}_smi.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
echo "== bench (b200 arm)"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-900
tail -3 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-900
echo "== C5 shape at 1e9 (two-field variant, 26 B) with the oracle check on a sample"
timeout 900 python tools/bench_c5.py --steps 10 --check 2>$OUT/${TAG}_c5.err | tee $OUT/${TAG}_c5.json | cut -c1-1500
tail -3 $OUT/${TAG}_c5.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
tail -9 $OUT/${TAG}_launches.csv | awk -F'","' '{print $5, $NF}'
echo "== ncu full capture of the express lane"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_sum_express -s 2 -c 1 -o $OUT/${TAG}_express \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
