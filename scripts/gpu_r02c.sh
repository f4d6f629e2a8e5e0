#!/bin/bash
# Round 2, GPU call C (1 GPU): full GPU test-suite, TMA ring variants for the SWAR kernel.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -60 | tee $OUT/${TAG}_pytest.log
echo "== variants (1e8 part)"
libs="libbydbgpu.so"
for f in skywalking-banyandb_b200/variants/*.so; do [ -e "$f" ] && libs="$libs variants/$(basename $f)"; done
timeout 900 python tools/time_variants.py $libs --steps 30 2>&1 | grep -v "^\s*$" | tail -20 | tee $OUT/${TAG}_variants.log
