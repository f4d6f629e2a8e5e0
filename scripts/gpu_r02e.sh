#!/bin/bash
# Round 2, GPU call E (1 GPU): tests, ring variants at 1e8 and 1e9, bench, ncu launch list + full capture of the express lane.
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== variants (1e8 part)"
libs="libbydbgpu.so"
for f in skywalking-banyandb_b200/variants/*.so; do [ -e "$f" ] && libs="$libs variants/$(basename $f)"; done
timeout 900 python tools/time_variants.py $libs --steps 30 2>&1 | grep -v "^\s*$" | tail -10 | tee $OUT/${TAG}_variants.log
BYDB_NO_EXPRESS=1 timeout 600 python tools/time_variants.py libbydbgpu.so --steps 30 2>&1 | grep -v "^\s*$" | tail -2 | sed 's/^/noexpress /' | tee -a $OUT/${TAG}_variants.log
echo "== bench (1e9)"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-1500
tail -3 $OUT/${TAG}_bench.err
echo "== 1e9 with ring variants / express off (resident only)"
for v in stage4k_s3 stage4k stages3; do
  BYDB_GPU_LIB=$PWD/skywalking-banyandb_b200/variants/$v.so timeout 900 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['scan_kernel_ms'])" | tee -a $OUT/${TAG}_variants.log
done
BYDB_NO_EXPRESS=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('noexpress', d['ms_per_step'], d['scan_kernel_ms'])" | tee -a $OUT/${TAG}_variants.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
tail -9 $OUT/${TAG}_launches.csv | awk -F'","' '{print $5, $NF}'
echo "== ncu full capture of the express lane"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_sum_express -s 2 -c 1 -o $OUT/${TAG}_express \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
