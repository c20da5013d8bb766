#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nmn_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
timeout 200 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|==|sum|epilogue" | grep -E "==|sum|mmas/tile=  72|mmas/tile= 576" | cut -c1-330 | tee gpurun_out/trace.txt
show() {
python - "$1" "$2" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print(sys.argv[2], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), {k: round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
}
for cfg in "0.7 0.8" "0.5 0.8" "0.6 0.7" "0.8 0.9" "0.7 1.0"; do
  set -- $cfg
  PNMN_CRIT=$1 PNMN_CRIT_BWD=$2 timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_crit.json 2> gpurun_out/bench_crit.err || tail -5 gpurun_out/bench_crit.err
  show gpurun_out/bench_crit.json "crit fwd=$1 bwd=$2"
done
