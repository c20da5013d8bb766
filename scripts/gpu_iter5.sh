#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print(sys.argv[2], 'n', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), {k: round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, {k: round(v,2) for k,v in d["host_ms_per_step"].items()}, 'e2e kernels', {k: round(v,2) for k,v in d['e2e'].get('kernel_ms_per_step',{}).items()})
PY
}
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_q$rep.json 2> gpurun_out/bench_q$rep.err || tail -5 gpurun_out/bench_q$rep.err
  show gpurun_out/bench_q$rep.json "run $rep"
done
