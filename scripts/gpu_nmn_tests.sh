#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/nmn_tests.log
: > $OUT
for impl in simt tc; do
  echo "##### PNMN_CONV_IMPL=$impl" >> $OUT
  PNMN_CONV_IMPL=$impl timeout 600 python -m pytest tests/test_nmn_gpu.py -q -s -x 2>&1 | grep -vE "^\s*$" | tail -40 >> $OUT
done
echo "##### wgrad kernel tests" >> $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -s -k "wgrad" 2>&1 | grep -E "rel err|passed|failed|Error" | tail -30 >> $OUT
tail -120 $OUT
