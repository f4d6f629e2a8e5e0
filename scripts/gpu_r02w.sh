#!/bin/bash
# Round 2, GPU call W (1 GPU): the round's closing run -- whole suite, smoke, both bench arms, launch list of the final code.
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $OUT/${TAG}_smi.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200 | tee $OUT/${TAG}_smoke.log
echo "== bench (b200 arm)"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'],'launches',j['gpu_launches'])
print('roofline',{k:j['roofline'][k] for k in ('achieved','peak','frac','frac_encoded','traffic')})
print('c2',j['c2_query']['scan_kernel_ms'],'keyed',j['stored_tag_group_by']['ms_per_step'])
e=j['e2e']; print('e2e ms',e['ms_per_step'], 'gather', e.get('unpinned_gather',{}).get('ms_per_step'))
print('cpu',j['cpu_baseline']['value'], j['cpu_baseline'].get('agrees_with_gpu'))
print('clocks',j['clocks'])
"
tail -2 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
tail -9 $OUT/${TAG}_launches.csv | awk -F'","' '{print $5, $NF}'
