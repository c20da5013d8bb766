#!/bin/bash
# experiment: stream priority of the generator's passes only (PNMN_JOINT_PRIORITY=2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for pr in ${PRIO_LIST:-0 2 0 2}; do
  PNMN_JOINT_PRIORITY=$pr timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_prio.json 2>gpurun_out/bench_prio.err || tail -5 gpurun_out/bench_prio.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_prio.json'))
print('priority $pr: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3))
PY
done
} 2>&1 | tee gpurun_out/prio_exp.txt
