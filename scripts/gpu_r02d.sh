#!/bin/bash
# Round 2, GPU call D (2 GPUs): the mailbox reduce across real devices -- pytest comm tests, the no-torch C ranks, bench at N=2.
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/${TAG}_smi.log
nvidia-smi topo -m 2>/dev/null | head -12 | tee -a $OUT/${TAG}_smi.log
echo "== comm tests on 2 devices"
timeout 900 python -m pytest tests -m gpu -q -k "scan_reduce or multi_process or partial_rows" 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -30 | tee $OUT/${TAG}_pytest.log
echo "== bench N=2 (strong scaling, 1e9 sharded over 2 ranks)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n2.err | tee $OUT/${TAG}_bench_n2.json | cut -c1-3000
tail -5 $OUT/${TAG}_bench_n2.err
