#!/bin/bash
# Round 2, GPU call L (1 GPU): A/B of the sparse masked decode against the serial decoder (2- and 3-stage rings) on the 1e8 part,
# then a full ncu capture (with source) of the fast lane on the masked query.
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
echo "== variants (1e8 part: masked = C2's query, allrows = sum+count over every row)"
timeout 900 python tools/time_variants.py libbydbgpu.so variants/serial3.so variants/serial2.so variants/sparse2.so --steps 30 2>&1 | grep -v "^$" | tee $OUT/${TAG}_variants.log
echo "== ncu full capture of the fast lane on the masked query"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_blocks_kernel -s 6 -c 1 -o $OUT/${TAG}_masked \
    python tools/time_variants.py libbydbgpu.so --steps 4 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log | cut -c1-200
