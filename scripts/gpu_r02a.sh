#!/bin/bash
# Round 2, GPU call A: every GPU test un-gated, the north-star bench (both arms), variant timing, ncu launch list + full capture.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee $OUT/${TAG}_smi.log
free -g | head -2 | tee -a $OUT/${TAG}_smi.log; nproc | tee -a $OUT/${TAG}_smi.log
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.log
echo "== pytest -m gpu (continue past first failure, names only)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed" | head -40 | tee $OUT/${TAG}_pytest_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== variants (1e8 part)"
bash scripts/gpu_variants_all.sh 2>&1 | tee $OUT/${TAG}_variants.log
echo "== bench (b200 arm, 1e9)"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-3000
tail -5 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-1500
echo "== ncu launch list (1e9, C3 query)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
tail -2 $OUT/${TAG}_ncu_launches.log | cut -c1-300
echo "== ncu full capture of scan_blocks_kernel (C3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_blocks -s 2 -c 1 -o $OUT/${TAG}_scan \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
ls -la $OUT | tail -20
