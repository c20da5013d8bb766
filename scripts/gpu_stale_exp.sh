#!/bin/bash
# experiment: what would the joint step's timeline be if the program compiler took no time (the module network runs an
# earlier step's sampled programs, compiled ahead), with and without an SM partition for the LSTM passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for stale in 0 1; do
for r in 0 32 48 64; do
  echo "== stale $stale reserve $r"
  PNMN_DIAG_STALE_PROGRAMS=$stale PNMN_JOINT_PRESTAGE=$((1-stale)) PNMN_JOINT_RESERVE_SMS=$r timeout 300 python scripts/joint_timeline.py 2>&1 | grep -v Warn | tr '\n' ';' | sed 's/ \+/ /g'
  echo
done
done
} | tee gpurun_out/stale_exp.txt
