#!/bin/bash
# usage: bash scripts/gpu_multi.sh N tag   (inside gpurun --gpus N)
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/${TAG}_bench_n$N.err | tee $OUT/${TAG}_bench_n$N.json
tail -5 $OUT/${TAG}_bench_n$N.err
