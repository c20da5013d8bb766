#!/bin/bash
# experiment: the sampled programs compiled as k independent plans on k host threads (PNMN_JOINT_CHUNKS) with 16 SMs reserved
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in ${CHUNK_CFGS:-"1 16" "2 16" "3 16" "4 16" "4 24" "6 16"}; do
  set -- $cfg
  PNMN_JOINT_CHUNKS=$1 PNMN_JOINT_RESERVE_SMS=$2 timeout 600 python bench.py --steps 80 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_chunks.json 2>gpurun_out/bench_chunks.err || tail -5 gpurun_out/bench_chunks.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunks.json'))
print('chunks $1 reserve $2: ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['host_ms_per_step'])
PY
done
} | tee gpurun_out/chunks_exp.txt
