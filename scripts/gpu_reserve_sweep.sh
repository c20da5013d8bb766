#!/bin/bash
# experiment: joint-step bench vs the number of SMs the module executor leaves to the LSTM passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for r in 0 16 24 32 40 0 32; do
  PNMN_JOINT_RESERVE_SMS=$r timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_res.json 2>gpurun_out/bench_res.err || tail -5 gpurun_out/bench_res.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_res.json'))
print('reserve $r: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
done
} | tee gpurun_out/reserve_sweep.txt
