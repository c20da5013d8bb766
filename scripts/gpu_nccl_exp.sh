#!/bin/bash
# experiment (N = 2): CTAs NCCL may use (NCCL_MAX_CTAS) next to the persistent executor kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
{
for cfg in ${CFGS:-"0 16" "4 16" "8 16" "16 16" "8 24"}; do
  set -- $cfg
  if [ "$1" = "0" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$1; fi
  if [ -n "$3" ]; then export NCCL_MIN_CTAS=$3; else unset NCCL_MIN_CTAS; fi
  PNMN_JOINT_RESERVE_SMS=$2 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_nccl.json 2> gpurun_out/bench_nccl.err || tail -20 gpurun_out/bench_nccl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_nccl.json').read().strip().splitlines()[-1])
print('N=$N NCCL_MAX_CTAS=$1 MIN=$3 reserve $2: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()})
PY
done
} 2>&1 | tee gpurun_out/nccl_exp_n$N.txt
