#!/bin/bash
# Round 2, GPU call M (1 GPU): sparse decode with rolled loops vs the serial decoder (A/B on the 1e8 part), the device page encoder
# (parity test + tools/bench_encode.py), the stored-tag group-by leg of the bench.
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
echo "== variants (1e8 part: masked = C2's query, allrows = sum+count over every row)"
timeout 900 python tools/time_variants.py libbydbgpu.so variants/serial3.so variants/serial2.so --steps 30 2>&1 | grep -v "^$" | tail -8 | tee $OUT/${TAG}_variants.log
echo "== encoder parity + keyed + C caller"
timeout 900 python -m pytest tests -m gpu -q -x -k "device_page_encoder or group_by_stored_tag or c_caller_on_the_device" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_new.log
echo "== encode bench"
timeout 900 python tools/bench_encode.py --values 50000000 2>$OUT/${TAG}_enc.err | tee $OUT/${TAG}_encode.json | cut -c1-1500
tail -3 $OUT/${TAG}_enc.err
echo "== ncu: fast lane on the masked query (raw page only)"
timeout 600 ncu --set full --clock-control none -k regex:scan_blocks_kernel -s 6 -c 1 -o $OUT/${TAG}_masked \
    python tools/time_variants.py libbydbgpu.so --steps 4 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
