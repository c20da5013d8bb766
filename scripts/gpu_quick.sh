#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_nmn_gpu.py -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|rel err" | cut -c1-220 | tail -14
for mode in split; do
PNMN_CLASSIFIER=$mode timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline frac', round(d['roofline']['frac'],4), d['classifier_math'])
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('host_ms_per_step'))
PY
done
