#!/bin/bash
# Round 2, GPU call K (1 GPU): sparse masked decode (delta_page_sparse) -- parity suite, bench (C3 value + C2 leg), C5 tool,
# compute-sanitizer memcheck over the smoke run and two parity tests.
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  |Error" | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'])
print('c2',json.dumps(j.get('c2_query'))[:600])
"
tail -3 $OUT/${TAG}_bench.err
echo "== C5"
timeout 900 python tools/bench_c5.py --steps 10 2>$OUT/${TAG}_c5.err | tee $OUT/${TAG}_c5.json | cut -c1-700
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_memcheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "groups_time_range_dict_pred or group_by_stored_tag_limits or dictionary_tag_shapes" > $OUT/${TAG}_memcheck_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_memcheck_tests.log
