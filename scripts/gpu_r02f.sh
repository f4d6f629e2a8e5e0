#!/bin/bash
# Round 2, GPU call F (2 GPUs): the comm tests across two devices (incl. the prepared collective), bench at N=2.
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench N=2 (strong scaling, 1e9 sharded over 2 ranks)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n2.err | grep "^{" | tee $OUT/${TAG}_bench_n2.json | cut -c1-1200
tail -5 $OUT/${TAG}_bench_n2.err
echo "== bench N=1"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench_n1.err | tee $OUT/${TAG}_bench_n1.json | cut -c1-1200
