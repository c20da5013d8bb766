#!/bin/bash
# experiment: the module network's backward pass + update issued with the NEXT step (PNMN_JOINT_DEFER_NMN=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
true
for cfg in ${CFGS:-"0 16" "1 16" "1 32" "1 48"}; do
  set -- $cfg
  PNMN_JOINT_DEFER_NMN=$1 PNMN_JOINT_RESERVE_SMS=$2 timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_defer.json 2>gpurun_out/bench_defer.err || tail -5 gpurun_out/bench_defer.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_defer.json'))
print('defer $1 reserve $2: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
done
} 2>&1 | tee gpurun_out/defer_exp.txt
