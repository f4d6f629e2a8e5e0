#!/bin/bash
# Round 2, GPU call R (2 GPUs): the collective with the final code -- comm tests (ranks on two devices), bench at N=2.
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
echo "== comm tests"
timeout 900 python -m pytest tests -m gpu -q -k "scan_reduce or multi_process or partial" 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -20 | tee $OUT/${TAG}_comm.log
echo "== bench N=2 (strong scaling, 1e9 sharded over 2 ranks)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n2.err | grep "^{" | tee $OUT/${TAG}_bench_n2.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'])
print('prepared',j.get('prepared_query'),'plain',j.get('plain_call',{}).get('ms_per_step'))
print('nccl ag',j.get('nccl_allgather_variant',{}).get('ms_per_step'),'ar',j.get('nccl_allreduce_variant',{}).get('ms_per_step'))
e=j['e2e']; print('e2e ms',e['ms_per_step'])
"
grep -E "Error|error" $OUT/${TAG}_bench_n2.err | head -5
