#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
for mode in 0 1; do
  PNMN_CLASSIFIER_TF32=$mode timeout 600 python bench.py --steps 20 --warmup 3 $([ $mode = 1 ] && echo --no-cpu-baseline) > gpurun_out/bench_tf32_$mode.json 2>gpurun_out/bench_tf32_$mode.err || tail -5 gpurun_out/bench_tf32_$mode.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_tf32_$mode.json'))
print('classifier_tf32=$mode value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline frac', round(d['roofline']['frac'],4), 'clocks', d['clocks'])
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('host_ms_per_step'))
print('pg', d.get('pg'), 'joint', d.get('joint'), 'cpu', d.get('cpu_baseline'))
PY
done
PNMN_CLASSIFIER_TF32=1 timeout 600 python -m pytest tests/test_nmn_gpu.py -q -s 2>&1 | grep -E "passed|failed|FAILED|rel err" | tail -12
