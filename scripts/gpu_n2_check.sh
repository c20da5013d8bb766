#!/bin/bash
# N = 2: the data-parallel GPU tests, then bench.py with 16 / 0 reserved SMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
{
true
for r in ${R_LIST:-16 0}; do
  PNMN_JOINT_RESERVE_SMS=$r timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_n${N}_r$r.json 2> gpurun_out/bench_n${N}_r$r.err || tail -20 gpurun_out/bench_n${N}_r$r.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_r$r.json').read().strip().splitlines()[-1])
print('N=$N reserve $r: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d["host_ms_per_step"].items()})
PY
done
} 2>&1 | tee gpurun_out/n${N}_check.txt
