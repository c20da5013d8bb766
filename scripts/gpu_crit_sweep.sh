#!/bin/bash
# experiment: joint-step bench vs the critical-sample thresholds of the program compiler (task granularity at 123 rows)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for c in "0.7 0.8" "0.5 0.6" "0.3 0.4" "0.0 0.0" "0.85 0.9"; do
  set -- $c
  PNMN_CRIT=$1 PNMN_CRIT_BWD=$2 timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_crit.json 2>gpurun_out/bench_crit.err || tail -5 gpurun_out/bench_crit.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_crit.json'))
print('crit $1 $2: ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, d['host_ms_per_step']['plan_create'])
PY
done
} | tee gpurun_out/crit_sweep.txt
