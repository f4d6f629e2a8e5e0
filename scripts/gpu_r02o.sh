#!/bin/bash
# Round 2, GPU call O (1 GPU): the shared-device collective after the page-locked allocations moved out of the query path --
# the mailbox test eight times in a row, the native multi-process test, then the whole suite.
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3 4 5 6 7 8; do
  timeout 400 python -m pytest tests -m gpu -q -x -k "scan_reduce_peer_mailboxes" 2>&1 | tail -1
done | tee $OUT/${TAG}_repeat.log
timeout 600 python -m pytest tests/test_c_abi_native.py -m gpu -q 2>&1 | tail -2
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|ERROR|passed|failed|^E  " | head -40 | tee $OUT/${TAG}_pytest.log
