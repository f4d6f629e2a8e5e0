#!/bin/bash
# First GPU call of a kernel-tuning session: time every prebuilt experiment variant (scripts/build_variants.sh, run on the
# build host beforehand) against the default library in one process, with result identity checks.
cd "$(dirname "$0")/.."
libs="libbydbgpu.so"
for f in skywalking-banyandb_b200/variants/*.so; do [ -e "$f" ] && libs="$libs variants/$(basename $f)"; done
timeout 900 python tools/time_variants.py $libs --steps 30 2>&1 | grep -v "^\s*$" | tail -40
