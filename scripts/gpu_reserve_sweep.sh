#!/bin/bash
# experiment: joint-step bench vs the number of SMs the module executor leaves to the LSTM passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for r in ${RESERVE_LIST:-0 8 12 16 20 0 8 12 16 20}; do
  PNMN_JOINT_RESERVE_SMS=$r timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_res.json 2>gpurun_out/bench_res.err || tail -5 gpurun_out/bench_res.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_res.json'))
print('reserve $r: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
done
} | tee gpurun_out/reserve_sweep2.txt
