#!/bin/bash
# experiment: joint-step timeline vs the number of SMs the module executor leaves to the LSTM passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seq2seq_gpu.py tests/test_kernels_gpu.py tests/test_nmn_gpu.py -q -x 2>&1 | tail -3
for prio in 0 1; do
for r in 0 32 64; do
  echo "== reserve $r priority $prio"
  PNMN_JOINT_PRIORITY=$prio PNMN_JOINT_RESERVE_SMS=$r timeout 300 python scripts/joint_timeline.py 2>&1 | grep -v Warn | tr '\n' ';' | sed 's/ \+/ /g'
  echo
done
done
timeout 300 python scripts/diag_pass.py 256 2>&1 | tail -4
