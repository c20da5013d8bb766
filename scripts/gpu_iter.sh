#!/bin/bash
# one development iteration on the GPU: NMN parity tests, executor trace, short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nmn_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -5
timeout 200 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|==|sum|epilogue" | grep -E "==|sum|mmas/tile=  72|mmas/tile= 576|epilogue" | cut -c1-330 | tee gpurun_out/trace.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline frac', round(d['roofline']['frac'],4))
print({k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}, {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
