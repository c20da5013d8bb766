#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps ${STEPS:-30} --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -20 gpurun_out/bench_n$N.err
python - gpurun_out/bench_n$N.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
