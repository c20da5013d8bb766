#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2>gpurun_out/bench2.err || tail -5 gpurun_out/bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['roofline']['launches_per_step'])
print(d['kernel_ms_per_step']); print(d['plan'])
PY
timeout 600 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum" | cut -c1-330 | tee gpurun_out/trace.txt
