#!/bin/bash
# task-granularity experiment: pairing / splitting thresholds of the program compiler
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "148 222" "100000 222" "100000 100000" "148 100000" "296 444" "74 111"; do
  set -- $cfg
  PNMN_PAIR_MIN=$1 PNMN_SPLIT_MAX=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err || tail -3 gpurun_out/b.err
  python - "$cfg" <<'PY'
import json, sys
d=json.load(open('gpurun_out/b.json'))
print('pair_min/split_max', sys.argv[1], '| ms/step', round(d['ms_per_step'],2), 'exec', round(d['kernel_ms_per_step']['conv_tc<2,2>'],3), 'plan', round(d['host_ms_per_step']['plan_create'],2))
PY
done
