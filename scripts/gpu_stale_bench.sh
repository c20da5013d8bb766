#!/bin/bash
# experiment: bench.py with the module network running an earlier step's sampled programs compiled ahead of time
# (upper bound of what a zero-latency program compiler would give), for a few SM partitions
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "0 16" "1 16" "1 32" "1 48" "1 0"; do
  set -- $cfg
  PNMN_DIAG_STALE_PROGRAMS=$1 PNMN_JOINT_PRESTAGE=$((1-$1)) PNMN_JOINT_RESERVE_SMS=$2 timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_stale.json 2>gpurun_out/bench_stale.err || tail -5 gpurun_out/bench_stale.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_stale.json'))
print('stale $1 reserve $2: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
done
} | tee gpurun_out/stale_bench.txt
