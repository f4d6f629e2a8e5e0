#!/bin/bash
# Round 2, GPU call G (2 GPUs): prepared collective with the capture trace, full GPU suite, bench at N=2 and N=1.
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
echo "== comm tests with BYDB_TRACE"
BYDB_TRACE=1 timeout 900 python -m pytest tests -m gpu -q -s -k "scan_reduce or multi_process" 2>&1 | grep -E "bydb\]|FAILED|ERROR|passed|failed|^E  " | grep -v "bydb cold" | head -30 | tee $OUT/${TAG}_comm.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
echo "== bench N=2 (strong scaling, 1e9 sharded over 2 ranks)"
BYDB_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n2.err | grep "^{" | tee $OUT/${TAG}_bench_n2.json | cut -c1-1000
grep -E "bydb\] prepared|Error|error" $OUT/${TAG}_bench_n2.err | grep -v "bydb cold" | head -10
echo "== bench N=1"
timeout 1500 python bench.py --steps 20 --warmup 3 2>$OUT/${TAG}_bench_n1.err | tee $OUT/${TAG}_bench_n1.json | cut -c1-1000
