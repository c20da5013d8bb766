#!/bin/bash
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-"X=1"}; do
  echo "### $v"
  env $v PNMN_E2E_TRACE=1 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "e2e step" | awk '{ if ($6+0 > 5.0) printf "%s%s(%s) ", $3, "", $6; n++; s+=$6 } END { printf "\n mean device period %.2f over %d steps\n", s/n, n }'
done
