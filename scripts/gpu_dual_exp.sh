#!/bin/bash
# dual MMA issuers: parity tests (guarded against hangs), then both bench workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -5
rc=${PIPESTATUS[0]}
timeout 600 python -m pytest tests/test_nmn_gpu.py -q -x 2>&1 | tail -5
for i in 1 2; do
  timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/bench_dual.json 2>gpurun_out/bench_dual.err || tail -5 gpurun_out/bench_dual.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_dual.json'))
x=d['extra']['executor']
print('joint ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],4), '| executor ms/step', round(x['ms_per_step'],3), 'frac', round(x['roofline']['frac'],4), 'exec ms', round(x['roofline']['kernel_ms_per_step'],3))
PY
done
} 2>&1 | tee gpurun_out/dual_exp.txt
