#!/bin/bash
# development iteration on the GPU: NMN parity tests, executor trace, short bench, end-to-end leg variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nmn_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -5
timeout 200 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|==|sum|epilogue" | grep -E "==|sum|mmas/tile=  72|mmas/tile= 576" | cut -c1-330 | tee gpurun_out/trace.txt
show() {
python - "$1" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4))
print('   ', {k: round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
}
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
show gpurun_out/bench_quick.json
for v in nocopy noread; do
  PNMN_E2E_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err || tail -5 gpurun_out/bench_$v.err
  show gpurun_out/bench_$v.json
done
PNMN_NO_PRECOMPILE=1 timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_nopre.json 2> gpurun_out/bench_nopre.err || tail -5 gpurun_out/bench_nopre.err
show gpurun_out/bench_nopre.json
