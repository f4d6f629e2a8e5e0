#!/bin/bash
# Round 2, GPU call S (8 GPUs): bench at N=8 and N=4 with the final code.
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
for N in 8 4; do
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 \
    2>$OUT/${TAG}_bench_n$N.err | grep "^{" | tee $OUT/${TAG}_bench_n$N.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'scan',j['scan_kernel_ms'],'device',j['device_ms_per_step'])
print('prepared',j.get('prepared_query'),'plain',j.get('plain_call',{}).get('ms_per_step'),'sustained',j.get('sustained',{}).get('ms_per_step'))
print('nccl ag',j.get('nccl_allgather_variant',{}).get('ms_per_step'),'ar',j.get('nccl_allreduce_variant',{}).get('ms_per_step'))
e=j['e2e']; print('e2e ms',e['ms_per_step'],'clocks',j.get('clocks'))
"
grep -E "Error|error" $OUT/${TAG}_bench_n$N.err | head -3
done
