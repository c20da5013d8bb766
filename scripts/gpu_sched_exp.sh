#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_exp.json 2>gpurun_out/bench_exp.err || tail -5 gpurun_out/bench_exp.err
  python - "$*" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_exp.json'))
print(sys.argv[1], '| value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'exec ms', round(d['kernel_ms_per_step']['conv_tc<2,2>'],2), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'], 'plan ms', round(d['host_ms_per_step']['plan_create'],2))
PY
}
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_nmn_gpu.py -q 2>&1 | tail -2
run A=1
run PNMN_LEVEL_ORDER=1
run PNMN_NOSPLIT=1
run PNMN_PAIR_ALWAYS=1
