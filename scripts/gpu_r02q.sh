#!/bin/bash
# Round 2, GPU call Q (1 GPU): whole suite (with the stored-tag sweep), bench with every leg (cold path with 128 index pieces).
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'])
e=j['e2e']; print('e2e ms',e['ms_per_step'], 'gather', e.get('unpinned_gather',{}).get('ms_per_step'))
print('traced',json.dumps(e.get('traced_step'))[:1500])
"
tail -3 $OUT/${TAG}_bench.err
