#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 8 32 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err || tail -5 gpurun_out/bench_c$c.err
  python - gpurun_out/bench_c$c.json $c <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print('connections', sys.argv[2], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
done
