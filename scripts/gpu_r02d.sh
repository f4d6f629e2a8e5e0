#!/bin/bash
# Round 2, GPU call D (2 GPUs): every GPU test (the comm tests then span two devices), ring variants, bench at N=1 and N=2.
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/${TAG}_smi.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== variants (1e8 part)"
libs="libbydbgpu.so"
for f in skywalking-banyandb_b200/variants/*.so; do [ -e "$f" ] && libs="$libs variants/$(basename $f)"; done
timeout 900 python tools/time_variants.py $libs --steps 30 2>&1 | grep -v "^\s*$" | tail -12 | tee $OUT/${TAG}_variants.log
echo "== express lane off (A/B)"
BYDB_NO_EXPRESS=1 timeout 600 python tools/time_variants.py libbydbgpu.so --steps 30 2>&1 | grep -v "^\s*$" | tail -2 | tee -a $OUT/${TAG}_variants.log
echo "== bench N=1"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench_n1.err | tee $OUT/${TAG}_bench_n1.json | cut -c1-1800
tail -3 $OUT/${TAG}_bench_n1.err
echo "== bench N=2 (strong scaling, 1e9 sharded over 2 ranks)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n2.err | tee $OUT/${TAG}_bench_n2.json | cut -c1-3500
tail -5 $OUT/${TAG}_bench_n2.err
