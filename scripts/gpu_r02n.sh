#!/bin/bash
# Round 2, GPU call N (1 GPU): default build (serial masked decode, 2-stage ring) -- full GPU suite, bench with every leg (stored-tag
# group-by, traced e2e step), racecheck on the smoke query.
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'])
print('c2',json.dumps(j.get('c2_query'))[:300])
print('keyed',json.dumps(j.get('stored_tag_group_by'))[:700])
e=j['e2e']; print('e2e ms',e['ms_per_step'], 'gather', e.get('unpinned_gather',{}).get('ms_per_step'))
print('traced',json.dumps(e.get('traced_step'))[:1800])
print('cpu',json.dumps(j.get('cpu_baseline'))[:300])
"
tail -3 $OUT/${TAG}_bench.err
echo "== racecheck (shared-memory hazards) on the smoke query"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/${TAG}_racecheck_smoke.log
