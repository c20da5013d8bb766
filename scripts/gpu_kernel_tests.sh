#!/bin/bash
# Runs the kernel-level GPU tests one pytest process per test id, each under its own timeout, so a
# hung kernel (e.g. an mbarrier that never completes) costs one test, not the whole gpurun call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/kernel_tests.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version --format=csv >> $OUT 2>&1
ids=$(python -m pytest tests/test_kernels_gpu.py --collect-only -q 2>/dev/null | grep "::")
for id in $ids; do
  echo "=== $id" >> $OUT
  timeout 120 python -m pytest "$id" -x -q -s 2>&1 | grep -E "rel err|abs err|passed|failed|Error|error|assert" | head -12 >> $OUT
  echo "exit=$?" >> $OUT
done
grep -c "passed" $OUT
tail -150 $OUT
