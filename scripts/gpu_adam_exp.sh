#!/bin/bash
# experiment: every model updated right behind its own backward pass (PNMN_JOINT_EARLY_ADAM=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
head -3 gpurun_out/adam_exp.txt 2>/dev/null | cut -c1-200
for ea in ${EA_LIST:-1 1 1 0}; do
  PNMN_JOINT_EARLY_ADAM=$ea timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_adam.json 2>gpurun_out/bench_adam.err || tail -5 gpurun_out/bench_adam.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_adam.json'))
print('early adam $ea: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['optimizer_launches_per_step'])
PY
done
} 2>&1 | tee gpurun_out/adam_exp.txt
