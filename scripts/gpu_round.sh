#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list and one full capture of the scan kernel.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== bench (b200 arm)"
timeout 900 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json
if [ "$2" != "noprof" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_launches.log 2>&1
tail -3 $OUT/${TAG}_ncu_launches.log
echo "== ncu full capture of scan_blocks_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_blocks -s 2 -c 1 -o $OUT/${TAG}_scan \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $OUT/${TAG}_smi.log
