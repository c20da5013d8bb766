#!/bin/bash
# experiment: executor task prefetch on (default) / off (PNMN_EXEC_DBG=8): parity tests, then both bench workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_nmn_gpu.py tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
for dbg in 8 0 8 0; do
  PNMN_EXEC_DBG=$dbg timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/bench_pf.json 2>gpurun_out/bench_pf.err || tail -5 gpurun_out/bench_pf.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_pf.json'))
x=d['extra']['executor']
print('dbg $dbg: joint ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, '| executor ms/step', round(x['ms_per_step'],3), 'frac', round(x['roofline']['frac'],4), 'exec ms', round(x['roofline']['kernel_ms_per_step'],3))
PY
done
} 2>&1 | tee gpurun_out/prefetch_exp.txt
