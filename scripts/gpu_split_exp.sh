#!/bin/bash
# split compile (forward-half plan + full plan on a helper thread): parity tests, then the bench with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_nmn_gpu.py tests/test_joint_gpu.py -q -x 2>&1 | tail -4
for sp in 1 0 1 0; do
  PNMN_SPLIT_COMPILE=$sp timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_split.json 2>gpurun_out/bench_split.err || tail -5 gpurun_out/bench_split.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_split.json'))
print('split $sp: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['host_ms_per_step'], 'frac', round(d['roofline']['frac'],3), d['parity_check'])
PY
done
} 2>&1 | tee gpurun_out/split_exp.txt
