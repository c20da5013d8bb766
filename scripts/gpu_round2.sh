#!/bin/bash
# end-of-round check: smoke, all GPU tests, the default bench line and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic')}, 'clocks', d['clocks'])
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('host_ms_per_step'), 'launches', d['gpu_launches'])
print('pg', d.get('pg'), '\njoint', d.get('joint'), '\ncpu', d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
