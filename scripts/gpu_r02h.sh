#!/bin/bash
# Round 2, GPU call H (8 GPUs): comm tests across devices, strong scaling of the 1e9 workload at N=8 and N=4.
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 | tee $OUT/${TAG}_smi.log
nproc | tee -a $OUT/${TAG}_smi.log
echo "== comm tests"
timeout 600 python -m pytest tests -m gpu -q -k "scan_reduce or multi_process" 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -20 | tee $OUT/${TAG}_comm.log
for N in 8 4; do
echo "== bench N=$N (strong scaling, 1e9 sharded over $N ranks)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n$N.err | grep "^{" | tee $OUT/${TAG}_bench_n$N.json | cut -c1-700
grep -E "Error|error" $OUT/${TAG}_bench_n$N.err | head -5
done
