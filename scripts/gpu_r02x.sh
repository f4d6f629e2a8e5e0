#!/bin/bash
# Round 2, GPU call X (1 GPU): SWAR sums under a row mask (delta_page_sum_masked) -- whole suite, A/B against the serial decoder
# on the 1e8 part, the C2 / C5 legs at 1e9.
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -30 | tee $OUT/${TAG}_pytest.log
echo "== variants"
timeout 600 python tools/time_variants.py libbydbgpu.so variants/serialsum.so --steps 30 2>&1 | grep -v "^$" | tail -4 | tee $OUT/${TAG}_variants.log
echo "== bench (C3 value + C2 leg)"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'])
print('c2',json.dumps(j.get('c2_query'))[:420])
print('keyed',j['stored_tag_group_by']['ms_per_step'])
"
echo "== C5"
timeout 600 python tools/bench_c5.py --steps 10 --check 2>$OUT/${TAG}_c5.err | tee $OUT/${TAG}_c5.json | cut -c1-300
