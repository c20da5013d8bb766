#!/bin/bash
# experiment: the reconstructor's backward pass + update deferred into the next step as well (PNMN_JOINT_DEFER_QR=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_joint_gpu.py -q -x -s -k "deferred" 2>&1 | grep -E "passed|failed|Error|differ" | tail -5
for q in 1 0; do
  PNMN_JOINT_DEFER_QR=$q timeout 400 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_dq.json 2>gpurun_out/bench_dq.err || tail -5 gpurun_out/bench_dq.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_dq.json'))
print('defer qr $q: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['parity_check']['max_abs_nmn_loss_err_vs_oracle'])
PY
done
} 2>&1 | tee gpurun_out/deferqr_exp.txt
