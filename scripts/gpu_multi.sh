#!/bin/bash
# usage: bash scripts/gpu_multi.sh N tag [c4]   (inside gpurun --gpus N): bench.py at N ranks, optionally the 1e9 strong-scaling config
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/${TAG}_bench_n$N.err | tee $OUT/${TAG}_bench_n$N.json
tail -5 $OUT/${TAG}_bench_n$N.err
if [ "$3" == "c4" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/bench_configs.py --steps 20 2>$OUT/${TAG}_c4_n$N.err | tee $OUT/${TAG}_c4_n$N.json
tail -3 $OUT/${TAG}_c4_n$N.err
fi
