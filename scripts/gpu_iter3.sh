#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() {
python - "$1" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4))
print('   ', {k: round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
}
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
show gpurun_out/bench_quick.json
