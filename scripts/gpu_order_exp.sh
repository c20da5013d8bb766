#!/bin/bash
# experiment: issue order of the backward passes in the fused joint step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for o in 0 1 2; do
  echo "== order $o"
  PNMN_JOINT_ORDER=$o timeout 300 python scripts/joint_timeline.py 2>&1 | grep -v Warn | tr '\n' ';' | sed 's/ \+/ /g'
  echo
  PNMN_JOINT_ORDER=$o timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bench_order.json 2>gpurun_out/bench_order.err || tail -5 gpurun_out/bench_order.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_order.json'))
print('order $o: ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})
PY
done
} | tee gpurun_out/order_exp.txt
