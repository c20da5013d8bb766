#!/bin/bash
# tests + bench (both classifier math modes) + executor timeline; everything lands in gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|rror" | tail -8
PNMN_CLASSIFIER_TF32=1 timeout 600 python -m pytest tests/test_nmn_gpu.py -q -s 2>&1 | grep -E "passed|failed|FAILED|rel err" | tail -12
for mode in 0 1; do
  PNMN_CLASSIFIER_TF32=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32_$mode.json 2>gpurun_out/bench_tf32_$mode.err || tail -5 gpurun_out/bench_tf32_$mode.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_tf32_$mode.json'))
print('classifier_tf32=$mode value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline frac', round(d['roofline']['frac'],4))
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('host_ms_per_step'))
PY
done
timeout 600 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum|busy" | cut -c1-330 > gpurun_out/trace.txt; head -12 gpurun_out/trace.txt | cut -c1-250
