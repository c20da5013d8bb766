#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/nmn_tests.log
: > $OUT
for impl in simt tc; do
  echo "##### PNMN_CONV_IMPL=$impl" >> $OUT
  PNMN_CONV_IMPL=$impl timeout 900 python -m pytest tests/test_nmn_gpu.py -q -s 2>&1 | grep -E "rel err|rel [0-9]|passed|failed|Error|assert|FAILED|Mismatch|Max |^E " | tail -60 >> $OUT
done
tail -150 $OUT
