#!/bin/bash
# Builds the kernel-experiment variants of libbydbgpu.so (all off in the default build) into skywalking-banyandb_b200/variants/.
# They travel to the GPU box with the snapshot; time them there in ONE call with
#   gpurun -- 'bash scripts/gpu_variants.sh libbydbgpu.so variants/dual.so variants/allrows.so ...'
set -e
cd "$(dirname "$0")/../skywalking-banyandb_b200"
mkdir -p variants
build() { echo "== $1: $2"; make -s variant OUT=variants/$1.so EXTRA="$2"; grep -A3 "scan_blocks_kernelILb1" build_variant.log | grep -E "registers|spill" || true; }
build earlystop "-DBYDB_EXP_EARLYSTOP"
build dual "-DBYDB_EXP_DUAL"
build dual_earlystop "-DBYDB_EXP_DUAL -DBYDB_EXP_EARLYSTOP"
build dual_allrows "-DBYDB_EXP_DUAL -DBYDB_EXP_ALLROWS"
build interior "-DBYDB_EXP_INTERIOR"
build dual_interior "-DBYDB_EXP_DUAL -DBYDB_EXP_INTERIOR"
build stages3 "-DBYDB_STAGES=3"
build allrows "-DBYDB_EXP_ALLROWS"
build es_int_all "-DBYDB_EXP_EARLYSTOP -DBYDB_EXP_INTERIOR -DBYDB_EXP_ALLROWS"
rm -f build_variant.log
ls -la variants
